import sys, time, json; sys.path.insert(0,'.')
import numpy as np
from automerge_classic_b200 import tracegen
from automerge_classic_b200.engine import GpuBackendDoc
for n in (100000, 1000000):
    t0 = time.time(); t = tracegen.generate('C3', n, 10); print('gen', n, time.time()-t0, len(t.blob), flush=True)
    for it in range(3):
        g = GpuBackendDoc()
        t0 = time.time(); fp = g.apply_packed_flat(t.blob, t.offsets, t.n_changes); dt = time.time()-t0
        print(n, 'apply wall %.1f ms' % (dt*1e3), 'ops/s %.3g' % (t.n_ops/dt), 'phases', [round(x,2) for x in g.timings()], 'launches', g.launches(), 'edits', len(fp.edits), flush=True)
        del g
