#!/bin/sh
# Round evidence on one GPU box (run under gpurun from the repository root): everything lands in gpurun_out/ and is copied
# into profiles/ by hand afterwards. Each step has its own timeout; a step that fails does not stop the rest.
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q > $O/ev_tests.log 2>&1
AMG_BENCH_WATCHDOG_S=400 timeout 500 python bench.py --steps 10 --warmup 3 > $O/ev_bench_default.json 2> $O/ev_bench_default.err
for W in C4 C2 C2b; do AMG_BENCH_WATCHDOG_S=300 timeout 330 python bench.py --workload $W --steps 5 --warmup 3 --no-extras > $O/ev_bench_$W.json 2> $O/ev_bench_$W.err; done
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/ev_bench_reference.json 2> $O/ev_bench_reference.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4500 --csv --log-file $O/ev_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/ev_ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_decode_tiles -s 58 -c 8 -f -o $O/ev_prof python tools/ab_decode.py automerge_classic_b200/libamgpu.so > $O/ev_ncu_full.log 2>&1
timeout 150 python tools/time_paths.py 200000 > $O/ev_time_paths.log 2>&1
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_engine_gpu.py -m gpu -q -k "decoded_rows or deflate_variants or unknown_columns or value_validation or utf16 or out_of_order or pointer_array or incremental_calls or load_without or duplicated" > $O/ev_memcheck.log 2>&1
tail -3 $O/ev_tests.log; tail -3 $O/ev_memcheck.log; ls -la $O/ev_*
