"""A/B timing of builds of libamgpu on one GPU: python tools/ab_decode.py libA.so libB.so ...
For each library: replay the C3 1M-op trace (3 calls, phase times of the last one) and time the decode kernels
(amg_bench_decode, CUDA events, 20 iterations). Development aid; `bench.py` is the measurement of record."""
import ctypes as C, hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from automerge_classic_b200 import tracegen
from automerge_classic_b200.engine import doc_class_for, _ErrStruct

if len(sys.argv) > 2:   # one process per library: builds of the same sources share GNU-unique symbols inside one process
    import subprocess
    for path in sys.argv[1:]:
        subprocess.call([sys.executable, os.path.abspath(__file__), path])
    sys.exit(0)
n = int(os.environ.get('AB_OPS', 1000000))
t = tracegen.generate('C3', n, 10)
for path in sys.argv[1:]:
    cls = doc_class_for(os.path.abspath(path))
    L = cls._library.L
    g = cls()
    err = _ErrStruct()
    L.amg_reserve(g.h, C.c_size_t(len(t.blob) + (1 << 20)), C.byref(err))
    for it in range(5):   # same document object, reset between calls: allocations are warm after the first call
        L.amg_reset(g.h, C.byref(err))
        fp = g.apply_packed_flat(t.blob, t.offsets, t.n_changes)
        ph = g.timings()
    digest = hashlib.sha1(fp.props.tobytes() + fp.edits.tobytes()).hexdigest()[:12]
    dev = sum(ph[1:12])
    ms_sha, ms_parse, ms_dec, algo = C.c_float(), C.c_float(), C.c_float(), C.c_uint64()
    rc = L.amg_bench_decode(g.h, 20, C.byref(ms_sha), C.byref(ms_parse), C.byref(ms_dec), C.byref(algo), C.byref(err))
    print('%-22s patch %s rc=%d sha %.3f parse %.3f decode %.3f ms | device ms/step %.3f -> %.4g ops/s | phases %s' % (
        os.path.basename(path), digest, rc, ms_sha.value, ms_parse.value, ms_dec.value, dev, t.n_ops / (dev / 1e3), [round(x, 3) for x in ph[:12]]), flush=True)
    del g
