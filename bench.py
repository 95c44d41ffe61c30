#!/usr/bin/env python
"""bench.py — ops/sec applied for automerge-classic's Backend.applyChanges path on B200.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json `metric`: "ops/sec applied (1M-op text trace)"): a makeText
change plus 10 actors x 100 000 single-op changes (70 % insert / 30 % delete, merge every 100 changes)
= 1 000 001 binary changes, ~132 MB, synthetic (seeded generator csrc/tracegen.cc). One *step* =
Backend.applyChanges(init(), all changes) -> incremental patch, on a document that is reset (not
re-allocated) between steps. With N GPUs every rank replays its own independent document (seed + rank,
config C5): no data-path collective, weak scaling; the time of a step is the max over ranks.

  value : ops/s with the change bytes already resident in HBM (CUDA events from after the upload to the end
          of the call: SHA-256, parse, gate, decode, op-set ordering, patch kernels, patch copy-out)
  e2e   : ops/s through the C ABI (amg_apply_changes_packed) from a pinned HOST buffer to the flat patch in
          host memory, host<->device copies inside the timed region (wall clock around the synchronous call)
  roofline: the column decode kernels (header parse + column expansion) re-run on resident data: algorithmic bytes
          of SURVEY.md §8d (encoded bytes + 48 B/op + 8 B/pred + 96 B/change) / CUDA-event time against the measured
          HBM peak; the SHA-256 kernel over the same bytes is ALU-bound and reported separately (`sha256_kernel`)
  cpu_baseline: the oracle (CPU restatement of the reference's algorithm, 1 core) on a bounded prefix
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
    os.environ['NCCL_DEBUG'] = 'WARN'   # keep stdout to the one JSON line (NCCL prints its version banner there)
# stdout carries exactly one JSON line: whatever native libraries print to file descriptor 1 (NCCL's version banner does,
# whatever NCCL_DEBUG says) goes to stderr instead; the JSON line is written through the saved descriptor
_JSON_OUT = os.fdopen(os.dup(1), 'w')
os.dup2(2, 1)


def _emit(line):
    print(line, file=_JSON_OUT, flush=True)


N_OPS, N_ACTORS = 1_000_000, 10
CPU_SAMPLE_OPS = 200_000


def read_traffic():
    """DRAM bytes per launch of the decode kernels from the committed ncu capture (None if the file is missing)."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic_r01.json')) as fh:
            return int(json.load(fh)['decode_total_bytes']), 'profiles/traffic_r01.json (ncu --set full, see profiles/README_r01.md)'
    except Exception:
        return None, None


def read_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], s[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons), 'samples': len(sm)}


def run_reference(args, rank, world):
    """The reference arm: the CPU restatement of the reference backend (oracle/, pinned on the reference's own
    tests) on the host cores. The reference is single-threaded JavaScript; the restatement is single-threaded."""
    if rank != 0:
        return
    import numpy as np
    import oracle
    from automerge_classic_b200 import tracegen
    oracle.build()
    t = tracegen.generate('C3', CPU_SAMPLE_OPS, N_ACTORS)
    times = []
    for i in range(args.warmup + args.steps):
        doc = oracle.OracleDoc()
        t0 = time.perf_counter()
        doc.apply_blob(t.blob.ctypes.data_as(C.c_void_p), t.offsets.ctypes.data_as(C.c_void_p), t.n_changes, want_patch=False)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    v = t.n_ops / (ms / 1e3)
    sample = 'first %d ops of the C3 trace (same generator, 10 actors), applyChanges(init(), all) on 1 core' % t.n_ops
    _emit(json.dumps({
        'impl': 'reference', 'metric': 'ops/sec applied (1M-op text trace)', 'value': v, 'unit': 'ops/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int64', 'data': 'synthetic',
        'config': {'workload': 'C3 text trace: 10 actors x 100k single-op changes (1M ops); reference arm runs a bounded prefix', 'sample_ops': t.n_ops},
        'cpu_baseline': {'value': v, 'unit': 'ops/s', 'cores': 1, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'ops/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='amgpu')
    ap.add_argument('--ops', type=int, default=N_OPS)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        return run_reference(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device — the engine has no CPU fallback')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from automerge_classic_b200 import build, tracegen
    if rank == 0:
        build.build_all()
    if world > 1:
        dist.barrier()
    from automerge_classic_b200.engine import GpuBackendDoc, default_library, _ErrStruct
    lib = default_library()
    L = lib.L

    trace = tracegen.generate('C3', args.ops, N_ACTORS, seed=tracegen.SEED + rank)
    nbytes = int(trace.offsets[-1])
    # the caller's buffer: pinned host memory (bench contract: inputs copied from pinned host memory every step)
    pinned = torch.empty(nbytes + 64, dtype=torch.uint8).pin_memory()
    pinned[:nbytes].copy_(torch.from_numpy(trace.blob))
    blob_ptr = C.c_void_p(pinned.data_ptr())
    offs = np.ascontiguousarray(trace.offsets)
    doc = GpuBackendDoc(device=local)
    err = _ErrStruct()
    L.amg_reserve(doc.h, C.c_size_t(nbytes + (1 << 20)), C.byref(err))

    def step():
        lib.check(L.amg_reset(doc.h, C.byref(err)), err)
        pp = C.c_void_p()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = L.amg_apply_changes_packed(doc.h, blob_ptr, offs.ctypes.data_as(C.c_void_p), C.c_size_t(trace.n_changes), 0, 1, C.byref(pp), C.byref(err))
        t_call = time.perf_counter() - t0
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        step.call_ms = t_call * 1e3
        lib.check(rc, err)
        n = C.c_size_t()
        L.amg_patch_bytes(pp, C.byref(n))
        L.amg_patch_free(pp)
        return dt, doc.timings(), n.value

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = doc.launches()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall, dev, patch_bytes = [], [], 0
    for _ in range(args.steps):
        dt, ph, patch_bytes = step()
        wall.append(dt)
        dev.append(sum(ph[1:12]) / 1e3)   # phases after the upload, CUDA events
        last_ph = ph
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    launches = (doc.launches() - launches0) // max(args.steps, 1)
    t_wall, t_dev = sum(wall) / len(wall), sum(dev) / len(dev)
    wall_steps, dev_steps = list(wall), list(dev)
    if world > 1:   # a step ends when the slowest rank is done
        tt = torch.tensor([t_wall, t_dev], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_wall, t_dev = float(tt[0]), float(tt[1])
    total_ops = trace.n_ops * world

    # decode roofline: re-run the decode kernels on the resident batch
    roofline = None
    if rank == 0:
        ms_sha, ms_parse, ms_dec, algo = C.c_float(), C.c_float(), C.c_float(), C.c_uint64()
        rc = L.amg_bench_decode(doc.h, 20, C.byref(ms_sha), C.byref(ms_parse), C.byref(ms_dec), C.byref(algo), C.byref(err))
        peak, peak_src = read_peaks()
        traffic, traffic_src = read_traffic()
        if rc == 0:
            # the HBM-bound part of the decode: header parse + column expansion. SHA-256 over the same bytes is
            # ALU-bound (64 rounds per 64-byte block) and is reported next to it, not folded into the HBM figure.
            t_dec = (ms_parse.value + ms_dec.value) / 1e3
            kernel_name = 'column decode = k_decode_tiles (fused header parse + column expansion, bulk-staged through shared memory) + DecodeColumnKernel for changes of more than 16 ops'
            ach = algo.value / t_dec / 1e9
            n_blocks = (trace.blob.size + 64 * trace.n_changes) / 64.0          # ~ message blocks incl. padding
            roofline = {'bound': 'hbm', 'kernel': kernel_name,
                        'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak, 'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
                        'algorithmic_bytes_per_launch': int(algo.value),
                        'ms': {'parse': ms_parse.value, 'decode_columns': ms_dec.value},
                        'sha256_kernel': {'bound': 'alu', 'ms': ms_sha.value, 'bytes_hashed': int(trace.blob.size),
                                          'gb_per_s': trace.blob.size / (ms_sha.value / 1e3) / 1e9 if ms_sha.value else None,
                                          'blocks_per_s': n_blocks / (ms_sha.value / 1e3) if ms_sha.value else None},
                        'with_sha256_gbs': algo.value / ((ms_sha.value + ms_parse.value + ms_dec.value) / 1e3) / 1e9}
        else:
            roofline = {'bound': 'hbm', 'achieved': None, 'peak': peak, 'unit': 'GB/s', 'frac': None, 'traffic': None, 'error': err.msg.decode()}

    # the other routes of SURVEY §8d, once each on rank 0 (wall clock through the C ABI): (ii) loadChanges + getPatch,
    # (iii) save, then load + getPatch of the saved document
    other = None
    if rank == 0:
        try:
            def wall(fn):
                torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, time.perf_counter() - t0
            lib.check(L.amg_reset(doc.h, C.byref(err)), err)
            _, t_lc = wall(lambda: doc.apply_packed_flat(blob_ptr, offs, trace.n_changes, want_patch=False))
            _, t_gp = wall(doc.get_patch_flat)
            saved, t_sv = wall(doc.save)
            d2, t_ld = wall(lambda: GpuBackendDoc(saved, device=local))
            _, t_gp2 = wall(d2.get_patch_flat)
            del d2
            other = {'loadChanges_plus_getPatch_ops_per_s': trace.n_ops / (t_lc + t_gp), 'loadChanges_ms': t_lc * 1e3, 'getPatch_ms': t_gp * 1e3,
                     'save_ms': t_sv * 1e3, 'saved_document_bytes': len(saved),
                     'load_plus_getPatch_ops_per_s': trace.n_ops / (t_ld + t_gp2), 'load_ms': t_ld * 1e3, 'getPatch_after_load_ms': t_gp2 * 1e3,
                     'note': 'single cold invocation each, host buffers in and out'}
        except Exception as e:   # never lose the headline line over the extras
            other = {'error': repr(e)[:200]}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        import oracle
        oracle.build()
        ts = tracegen.generate('C3', CPU_SAMPLE_OPS, N_ACTORS)
        od = oracle.OracleDoc()
        t0 = time.perf_counter()
        od.apply_blob(ts.blob.ctypes.data_as(C.c_void_p), ts.offsets.ctypes.data_as(C.c_void_p), ts.n_changes, want_patch=False)
        dt = time.perf_counter() - t0
        cpu = {'value': ts.n_ops / dt, 'unit': 'ops/s', 'cores': 1, 'kind': 'port',
               'sample': 'first %d ops of the same C3 workload, oracle restatement of backend/new.js (not V8), %.1f s' % (ts.n_ops, dt)}

    if rank == 0 and os.environ.get('AMG_BENCH_MARKS'):
        buf = C.create_string_buffer(4096)
        L.amg_debug_marks(doc.h, buf, 4096)
        print('marks:', buf.value.decode(), file=sys.stderr)
    if rank == 0:
        clocks = sampler.summary()
        _emit(json.dumps({
            'metric': 'ops/sec applied (1M-op text trace)', 'value': total_ops / t_dev, 'unit': 'ops/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': t_wall * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int64', 'data': 'synthetic',
            'config': {'workload': 'C3 text trace: makeText + 10 actors x 100k single-op changes, 70% insert / 30% delete (SURVEY.md 8d); one independent document per GPU (C5)',
                       'ops_per_gpu': trace.n_ops, 'changes_per_gpu': trace.n_changes, 'change_bytes_per_gpu': nbytes, 'parallelism': 'replicas x%d' % world,
                       'l2': 'inputs (%.0f MB) + working tables exceed the 126 MB L2; document reset every step' % (nbytes / 1e6),
                       'device_ms_per_step': t_dev * 1e3, 'wall_ms_steps': [round(x * 1e3, 3) for x in wall_steps], 'call_return_ms_last_step': round(step.call_ms, 3), 'abi_call_ms_last_step': round(last_ph[23], 3), 'device_ms_steps': [round(x * 1e3, 3) for x in dev_steps],
                       'phase_ms_last_step': dict(zip(['stage_upload', 'sha256', 'parse_gate', 'actors_decode', 'opset', 'patch_groups_props', 'patch_list_index', 'patch_edits_copyout', 'heads_commit'], [round(x, 3) for x in last_ph[:9]])),
                       'host_marks_ms': [round(x, 3) for x in last_ph[12:22]], 'other_paths': other},
            'e2e': {'value': total_ops / t_wall, 'unit': 'ops/s', 'h2d_bytes_per_step': nbytes + 8 * (trace.n_changes + 1), 'd2h_bytes_per_step': patch_bytes},
            'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu, 'clocks': clocks}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
