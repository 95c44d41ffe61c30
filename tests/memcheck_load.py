"""Small load / save / column-decoder workload for `compute-sanitizer --tool memcheck python tests/memcheck_load.py`
(the pytest GPU suite is too long under the sanitizer). Every document takes the parallel column decoders."""
import os, sys
os.environ['AMG_PAR_DOC_MIN'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import oracle, parity_checks
from automerge_classic_b200.engine import GpuBackendDoc
for cfg, n, a in [('C3', 3000, 3), ('C4', 1500, 4), ('C8', 300, 3)]:
    parity_checks.check_load(GpuBackendDoc, oracle, cfg, n, a)
parity_checks.check_save_after_load(GpuBackendDoc, oracle, 'C6', 200, 1)
parity_checks.check_rust_document(GpuBackendDoc)
print(parity_checks.check_column_decoders(GpuBackendDoc, 7, 6))
print('memcheck workload ok')
