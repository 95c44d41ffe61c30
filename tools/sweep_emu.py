"""Randomized parity sweep on the emulation build (no GPU): random trace configurations (C3 - C8, 1 - 6 actors, 40 - 600 ops,
call sizes 1 - 1000; SWEEP_MINOPS / SWEEP_MAXOPS / SWEEP_CONTINUE in the environment change that) replayed call by call through tests/_emu/libamgpu_emu.so and the oracle; every incremental patch, the final
getPatch, the op table with its succ lists, decoded rows and save() must be identical.  python tools/sweep_emu.py <cases> [seed]
Development aid (the kernels' functors run as serial loops here); the parity tests proper are `pytest -m gpu`."""
import os, random, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import parity_checks, oracle
from automerge_classic_b200.engine import doc_class_for
subprocess.check_call([os.path.join(ROOT, 'tests', '_emu', 'build.sh')])
oracle.build()
Doc = doc_class_for(os.path.join(ROOT, 'tests', '_emu', 'libamgpu_emu.so'))
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 20260923)
t0 = time.time(); done = skipped = 0; mismatches = []
for k in range(cases):
    cfg = rnd.choice(['C3', 'C4', 'C6', 'C6', 'C7', 'C8', 'C8'])
    n, a, chunk, seed = rnd.randrange(int(os.environ.get('SWEEP_MINOPS', 40)), int(os.environ.get('SWEEP_MAXOPS', 600))), rnd.randrange(1, 7), rnd.choice([1, 2, 3, 7, 20, 97, 1000]), rnd.randrange(1 << 30)
    try:
        if cfg in ('C6', 'C8'):
            ok = parity_checks.check_rich_list(Doc, oracle, seed, n, a, chunk, cfg=cfg)
            skipped += 0 if ok else 1
        elif cfg == 'C7':
            parity_checks.check_counters(Doc, oracle, seed, n, a, chunk)
        else:
            parity_checks.check_trace_parity(Doc, oracle, cfg, n, a)
            parity_checks.check_decoded_rows_trace(Doc, oracle, cfg, n, a)
        if k % 5 == 0 and cfg != 'C8':
            parity_checks.check_save(Doc, oracle, cfg, n, a, chunk=max(chunk, 5))
    except AssertionError as e:
        print('MISMATCH', cfg, n, a, chunk, seed, str(e)[:400], flush=True)
        if not os.environ.get('SWEEP_CONTINUE'):
            sys.exit(1)
        mismatches.append((cfg, n, a, chunk, seed)); continue
    done += 1
    if done % 20 == 0:
        print('%d cases, %d without a reference answer (block-boundary non-termination), %.0f s' % (done, skipped, time.time() - t0), flush=True)
print('sweep: %d cases identical to the oracle (%d had no reference answer), %d mismatches %s, %.0f s' % (done, skipped, len(mismatches), mismatches[:8], time.time() - t0))
