"""CPU-side checks: the C-ABI library exports every symbol include/amgpu.h declares (no compute without a
GPU), fails loudly without a device, the trace generator's bytes are valid for the oracle, and the
multi-rank bench logic (one independent document per rank, max-over-ranks step time) under gloo."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built():
    from automerge_classic_b200 import build
    return build.build_all()


def test_header_symbols_exported(built):
    header = open(os.path.join(ROOT, 'include', 'amgpu.h')).read()
    names = sorted(set(re.findall(r'\b(amg_[a-z_0-9]+)\s*\(', header)))
    assert len(names) >= 25
    lib = C.CDLL(built[0])
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from automerge_classic_b200 import Backend
    from automerge_classic_b200.engine import AmgError
    with pytest.raises(AmgError, match='no CUDA device|CUDA'):
        Backend.init()


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'automerge_classic_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.cc', '.h')):
                src = open(os.path.join(dirpath, f), errors='replace').read()
                assert 'import oracle' not in src and 'from oracle' not in src and 'liboracle' not in src and '#include "../../oracle' not in src, f


@pytest.mark.parametrize('cfg,n,a', [('C1', 0, 0), ('C2', 500, 0), ('C2b', 2000, 0), ('C3', 3000, 10), ('C4', 3000, 5)])
def test_tracegen_bytes_are_valid_changes(oracle_mod, built, cfg, n, a):
    """The generator has its own encoder; the oracle (pinned on the reference) must accept every change, and the
    host-side Python mirror must re-encode each decoded change to identical bytes (uncompressed form)."""
    from automerge_classic_b200 import tracegen, columnar
    t = tracegen.generate(cfg, n, a)
    ch = t.changes()
    doc = oracle_mod.OracleDoc()
    patch = doc.apply_changes(ch)
    assert patch['pendingChanges'] == 0 and patch['maxOp'] >= t.n_ops // max(a, 1)
    assert sum(len(oracle_mod.decode_change(c)['ops']) for c in ch[:50]) == sum(len(columnar.decode_change(c)['ops']) for c in ch[:50])
    for c in ch[:40]:
        d = columnar.decode_change(c)
        assert columnar.encode_change(d, compress=False) == columnar.inflate_change(c)
    # determinism
    assert np.array_equal(tracegen.generate(cfg, n, a).blob, t.blob)


GLOO_SCRIPT = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
from automerge_classic_b200 import tracegen
import oracle
t = tracegen.generate('C3', 600, 3, seed=tracegen.SEED + rank)          # one independent document per rank (config C5)
doc = oracle.OracleDoc(); p = doc.apply_changes(t.changes())
mine = torch.tensor([float(rank + 1), float(t.n_ops)], dtype=torch.float64)   # (pretend step time, ops)
tmax = mine.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
tsum = mine.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
heads = [None] * world; dist.all_gather_object(heads, p['deps'])
if rank == 0:
    print(json.dumps({'step_time': float(tmax[0]), 'total_ops': float(tsum[1]), 'distinct_docs': len({tuple(h) for h in heads})}))
dist.destroy_process_group()
'''


def test_two_rank_gloo_replicas(built, oracle_mod, tmp_path):
    script = tmp_path / 'gloo_replicas.py'
    script.write_text(GLOO_SCRIPT % {'root': ROOT})
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
                          '--master-port', '29517', str(script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith('{')][-1])
    assert res['step_time'] == 2.0 and res['total_ops'] == 2 * 601 and res['distinct_docs'] == 2
