"""Synthetic edit traces in automerge-classic's binary change format (csrc/tracegen.cc)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
CONFIGS = {'C1': 1, 'C2': 2, 'C2b': 22, 'C3': 3, 'C4': 4, 'C6': 6, 'C7': 7, 'C8': 8}
SEED = 0xA17E0C1A551C


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'libamgtrace.so')
        if not os.path.exists(path):
            from .build import build_tracegen
            build_tracegen()
        _LIB = C.CDLL(path)
        _LIB.amg_trace_free.argtypes = [C.c_void_p]
    return _LIB


class Trace:
    """blob: uint8 numpy array; offsets: uint64 numpy array of n_changes + 1 entries."""

    def __init__(self, blob, offsets, n_ops, name):
        self.blob, self.offsets, self.n_ops, self.name = blob, offsets, n_ops, name

    @property
    def n_changes(self):
        return len(self.offsets) - 1

    def changes(self):
        b = self.blob.tobytes()
        return [b[int(self.offsets[i]):int(self.offsets[i + 1])] for i in range(self.n_changes)]


def generate(config, n_ops=0, n_actors=0, seed=SEED):
    L = _lib()
    blob, blen, offs, n, total = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t(), C.c_uint64()
    rc = L.amg_trace_generate(CONFIGS[config], C.c_uint64(seed), C.c_uint64(n_ops), int(n_actors), C.byref(blob), C.byref(blen), C.byref(offs), C.byref(n), C.byref(total))
    if rc != 0:
        raise ValueError('unknown trace config %r' % (config,))
    b = np.frombuffer(C.string_at(blob, blen.value), dtype=np.uint8).copy()
    o = np.frombuffer(C.string_at(offs, 8 * (n.value + 1)), dtype=np.uint64).copy()
    L.amg_trace_free(blob)
    L.amg_trace_free(offs)
    return Trace(b, o, int(total.value), config)
