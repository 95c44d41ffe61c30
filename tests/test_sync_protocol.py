"""Sync protocol functions (automerge_classic_b200/sync.py) over the engine: CPU run on the serial emulation build, GPU run
on libamgpu.so. Scenarios follow the reference's test/sync_test.js (which needs the JavaScript frontend to run as is)."""
import os
import subprocess

import pytest

import parity_checks

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def emu_doc():
    subprocess.check_call([os.path.join(HERE, '_emu', 'build.sh')])
    from automerge_classic_b200 import build
    build.build_tracegen()
    from automerge_classic_b200.engine import doc_class_for
    return doc_class_for(os.path.join(HERE, '_emu', 'libamgpu_emu.so'))


def test_sync_protocol_emu(emu_doc):
    assert parity_checks.check_sync_protocol(emu_doc)


@pytest.mark.parametrize('seed', [1, 2, 3, 4, 5, 6])
def test_sync_random_emu(emu_doc, seed):
    assert parity_checks.check_sync_random(emu_doc, seed) > 0


@pytest.mark.parametrize('seed', [1, 2, 3, 4, 5, 6, 10, 11])
def test_sync_transcripts_equal_emu(emu_doc, oracle_mod, seed):
    assert parity_checks.check_sync_transcripts_equal(emu_doc, oracle_mod, seed) > 0


def test_get_changes_differential_emu(emu_doc, oracle_mod):
    assert parity_checks.check_get_changes_differential(emu_doc, oracle_mod) > 0
    assert parity_checks.check_graph_queries_differential(emu_doc, oracle_mod) > 0


@pytest.mark.gpu
def test_sync_protocol_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from automerge_classic_b200 import build
    build.build_all()
    from automerge_classic_b200.engine import GpuBackendDoc
    assert parity_checks.check_sync_protocol(GpuBackendDoc)


@pytest.mark.gpu
@pytest.mark.parametrize('seed', [1, 2, 3])
def test_sync_random_gpu(seed):
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from automerge_classic_b200.engine import GpuBackendDoc
    assert parity_checks.check_sync_random(GpuBackendDoc, seed) > 0


@pytest.mark.gpu
def test_get_changes_and_transcripts_gpu(oracle_mod):
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from automerge_classic_b200.engine import GpuBackendDoc
    assert parity_checks.check_get_changes_differential(GpuBackendDoc, oracle_mod, queries=60) > 0
    assert parity_checks.check_graph_queries_differential(GpuBackendDoc, oracle_mod, rounds=6) > 0
    for seed in (5, 10):
        assert parity_checks.check_sync_transcripts_equal(GpuBackendDoc, oracle_mod, seed) > 0
