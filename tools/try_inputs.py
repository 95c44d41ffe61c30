"""Development aid: one applyChanges call per kind of input buffer (pageable, pinned, device) on a small C3 trace."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from automerge_classic_b200 import tracegen
from automerge_classic_b200.engine import GpuBackendDoc, default_library, _ErrStruct
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
kinds = sys.argv[2:] or ['pageable', 'pinned', 'device']
t = tracegen.generate('C3', n, 10)
nbytes = int(t.offsets[-1])
lib = default_library(); L = lib.L
pinned = torch.empty(nbytes + 64, dtype=torch.uint8).pin_memory(); pinned[:nbytes].copy_(torch.from_numpy(t.blob))
dev = pinned.to('cuda:0'); torch.cuda.synchronize()
offs = np.ascontiguousarray(t.offsets)
for kind in kinds:
    doc, err = GpuBackendDoc(), _ErrStruct()
    ptr = {'pageable': t.blob.ctypes.data, 'pinned': pinned.data_ptr(), 'device': dev.data_ptr()}[kind]
    for it in range(3):
        L.amg_reset(doc.h, C.byref(err))
        pp = C.c_void_p(); t0 = time.perf_counter()
        print(kind, 'call', it, flush=True)
        rc = L.amg_apply_changes_packed(doc.h, C.c_void_p(ptr), offs.ctypes.data_as(C.c_void_p), C.c_size_t(t.n_changes), 0, 1, C.byref(pp), C.byref(err))
        print(kind, 'rc', rc, err.msg.decode() if rc else '', '%.2f ms' % ((time.perf_counter() - t0) * 1e3), [round(x, 2) for x in doc.timings()[:9]], flush=True)
        L.amg_patch_free(pp)
