"""Pins the CPU oracle on the reference's own tests: every assertion of test/new_backend_test.js
(byte-exact doc columns, block metadata, Bloom bits, patches, error messages) and of
test/backend_test.js (patches through the Backend facade, save/load, hash-graph queries), replayed
from the fixtures that tests/jsfixtures/extract.py wrote to tests/golden/."""
import pytest

import replay
from automerge_classic_b200.backend import RangeError as FacadeRangeError


def _cases(name):
    return [pytest.param(t, id=t['name'][-70:]) for t in replay.load(name) if 'skipped' not in t]


@pytest.fixture(scope='module')
def replayer(oracle_mod):
    return replay.Replayer(oracle_mod.OracleDoc, (oracle_mod.OracleError, ValueError, TypeError, RuntimeError, FacadeRangeError), structural=True)


@pytest.mark.parametrize('test', _cases('new_backend_test.json'))
def test_new_backend(replayer, test):
    fails = replayer.run_test(test)
    assert not fails, '\n'.join(fails[:5])


@pytest.mark.parametrize('test', _cases('backend_test.json'))
def test_backend(replayer, test):
    fails = replayer.run_test(test)
    assert not fails, '\n'.join(fails[:5])


def test_fixture_counts():
    nb, bt = replay.load('new_backend_test.json'), replay.load('backend_test.json')
    assert len(nb) == 40 and not [t for t in nb if 'skipped' in t]
    assert len(bt) == 59 and len([t for t in bt if 'skipped' in t]) <= 1
    n_asserts = sum(1 for t in nb + bt for s in t['steps'] if s['op'].startswith('assert') or s['op'] == 'check_columns')
    assert n_asserts > 1200
