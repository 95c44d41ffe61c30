"""BASELINE.json's full size against the oracle: the document the CUDA engine reaches for the 1M-op C3 trace has the save()
digest and heads the oracle produced (tests/golden/full_size_c3.json; the oracle needs 75 s for it, so its result is
committed, not recomputed). Runs last: the file name sorts after the other test modules."""
import pytest

import parity_checks


@pytest.mark.gpu
def test_full_size_oracle_fingerprint():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from automerge_classic_b200 import build
    build.build_all()
    from automerge_classic_b200.engine import GpuBackendDoc
    parity_checks.check_full_size_properties(GpuBackendDoc, golden=True)


@pytest.mark.gpu
@pytest.mark.parametrize('cfg', ['C2', 'C2b', 'C4_100k', 'C3'])
def test_full_size_patch_and_document(cfg):
    """Incremental patch, getPatch, save() and heads at BASELINE.json's full size of every workload against
    the oracle's committed digests (tests/golden/full_size.json)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from automerge_classic_b200 import build
    build.build_all()
    from automerge_classic_b200.engine import GpuBackendDoc
    parity_checks.check_full_size_fingerprint(GpuBackendDoc, cfg)
