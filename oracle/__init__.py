"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes wrapper around oracle/liboracle.so, the CPU restatement of automerge-classic's
Backend.applyChanges / load / getPatch path (see codec.hpp, columnar.hpp, opset.hpp, backend.hpp
for the reference file:line each function follows).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package. The product package never does.
"""
import ctypes as C
import json
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, 'liboracle.so')
    srcs = [os.path.join(_HERE, f) for f in ('capi.cc', 'backend.hpp', 'opset.hpp', 'columnar.hpp', 'codec.hpp')]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, 'liboracle.so')
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_new.restype = C.c_void_p
        _LIB.orc_clone.restype = C.c_void_p
        _LIB.orc_clone.argtypes = [C.c_void_p]
        _LIB.orc_free.argtypes = [C.c_void_p]
        _LIB.orc_free_mem.argtypes = [C.c_void_p]
    return _LIB


class OracleError(Exception):
    """Carries the reference's error class name ('RangeError' / 'TypeError') and message."""

    def __init__(self, text):
        super().__init__(text)
        self.kind, _, self.message = text.partition(': ')


def _check(rc, err):
    if rc != 0:
        msg = C.string_at(err).decode('utf-8', 'replace') if err else 'unknown error'
        if err:
            lib().orc_free_mem(err)
        raise OracleError(msg)


def _take_str(p):
    s = C.string_at(p).decode('utf-8', 'replace')   # (corrupted test inputs may carry bytes that are not UTF-8)
    lib().orc_free_mem(p)
    return s


def _blob(changes):
    offs = (C.c_uint64 * (len(changes) + 1))()
    total = 0
    for i, c in enumerate(changes):
        offs[i] = total
        total += len(c)
    offs[len(changes)] = total
    data = b''.join(bytes(c) for c in changes)
    return data, offs


def _unpack(ptr, length):
    raw = C.string_at(ptr, length)
    lib().orc_free_mem(ptr)
    import struct
    n = struct.unpack_from('<Q', raw, 0)[0]
    offs = struct.unpack_from('<%dQ' % (n + 1), raw, 8)
    base = 8 * (n + 2)
    return [raw[base + offs[i]: base + offs[i + 1]] for i in range(n)]


class OracleDoc:
    """Mirror of the reference's BackendDoc (backend/new.js:1694)."""

    def __init__(self, data=None, _handle=None):
        L = lib()
        if _handle is not None:
            self.h = _handle
        elif data is None:
            self.h = C.c_void_p(L.orc_new())
        else:
            out, err = C.c_void_p(), C.c_char_p()
            data = bytes(data)
            _check(L.orc_load(data, C.c_size_t(len(data)), C.byref(out), C.byref(err)), err.value and C.cast(err, C.c_void_p))
            self.h = out

    def __del__(self):
        try:
            if self.h:
                lib().orc_free(self.h)
                self.h = None
        except Exception:
            pass

    def clone(self):
        return OracleDoc(_handle=C.c_void_p(lib().orc_clone(self.h)))

    def apply_changes(self, changes, is_local=False, want_patch=True):
        data, offs = _blob(changes)
        out, err = C.c_void_p(), C.c_void_p()
        rc = lib().orc_apply_changes(self.h, data, offs, C.c_size_t(len(changes)), int(is_local),
                                     C.byref(out) if want_patch else None, C.byref(err))
        _check(rc, err.value)
        return json.loads(_take_str(out)) if want_patch else None

    def apply_blob(self, data, offs_u64, n, want_patch=False):
        """Zero-copy variant for timing: data is bytes-like, offs_u64 a ctypes/numpy uint64 array of n+1 offsets."""
        out, err = C.c_void_p(), C.c_void_p()
        rc = lib().orc_apply_changes(self.h, data, offs_u64, C.c_size_t(n), 0, C.byref(out) if want_patch else None, C.byref(err))
        _check(rc, err.value)
        return json.loads(_take_str(out)) if want_patch else None

    def get_patch(self):
        out, err = C.c_void_p(), C.c_void_p()
        _check(lib().orc_get_patch(self.h, C.byref(out), C.byref(err)), err.value)
        return json.loads(_take_str(out))

    def save(self):
        out, n, err = C.c_void_p(), C.c_size_t(), C.c_void_p()
        _check(lib().orc_save(self.h, C.byref(out), C.byref(n), C.byref(err)), err.value)
        b = C.string_at(out, n.value)
        lib().orc_free_mem(out)
        return b

    def heads(self):
        out, err = C.c_void_p(), C.c_void_p()
        _check(lib().orc_get_heads(self.h, C.byref(out), C.byref(err)), err.value)
        return json.loads(_take_str(out))

    def clock(self):
        out, err = C.c_void_p(), C.c_void_p()
        _check(lib().orc_clock_json(self.h, C.byref(out), C.byref(err)), err.value)
        return json.loads(_take_str(out))

    def max_op(self):
        v = C.c_int64()
        lib().orc_max_op(self.h, C.byref(v))
        return v.value

    def hash_by_actor(self, actor, index):
        out, err = C.c_void_p(), C.c_void_p()
        _check(lib().orc_hash_by_actor(self.h, actor.encode(), C.c_int64(index), C.byref(out), C.byref(err)), err.value)
        return _take_str(out) or None

    def get_change_by_hash(self, hash_):
        out, n, err = C.c_void_p(), C.c_size_t(), C.c_void_p()
        _check(lib().orc_get_change_by_hash(self.h, hash_.encode(), C.byref(out), C.byref(n), C.byref(err)), err.value)
        if not out.value:
            return None
        b = C.string_at(out, n.value)
        lib().orc_free_mem(out)
        return b

    def get_changes(self, have_deps):
        out, n, err = C.c_void_p(), C.c_size_t(), C.c_void_p()
        _check(lib().orc_get_changes(self.h, ''.join(have_deps).encode(), C.byref(out), C.byref(n), C.byref(err)), err.value)
        return _unpack(out, n.value)

    def get_changes_added(self, old):
        out, n, err = C.c_void_p(), C.c_size_t(), C.c_void_p()
        _check(lib().orc_get_changes_added(self.h, old.h, C.byref(out), C.byref(n), C.byref(err)), err.value)
        return _unpack(out, n.value)

    def get_missing_deps(self, heads=()):
        out, err = C.c_void_p(), C.c_void_p()
        _check(lib().orc_get_missing_deps(self.h, ''.join(heads).encode(), C.byref(out), C.byref(err)), err.value)
        return json.loads(_take_str(out))

    def blocks(self):
        out, err = C.c_void_p(), C.c_void_p()
        _check(lib().orc_blocks_json(self.h, C.byref(out), C.byref(err)), err.value)
        blocks = json.loads(_take_str(out))
        for b in blocks:
            b['columns'] = {int(k): bytes.fromhex(v) for k, v in b['columns'].items()}
            b['bloom'] = bytes.fromhex(b['bloom'])
            for k, v in list(b.items()):
                if v == 'undefined' or v == '\x00undefined':
                    b[k] = Undefined
        return blocks

    def dump_ops(self):
        """Doc-ordered op table: (rows[n,12] int64, succ[m,2] int64, actorIds)."""
        import numpy as np
        rows, n, succ, m, actors, err = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_void_p()
        _check(lib().orc_dump_ops(self.h, C.byref(rows), C.byref(n), C.byref(succ), C.byref(m), C.byref(actors), C.byref(err)), err.value)
        r = np.frombuffer(C.string_at(rows, 8 * 12 * n.value), dtype=np.int64).reshape(-1, 12).copy()
        s = np.frombuffer(C.string_at(succ, 8 * 2 * m.value), dtype=np.int64).reshape(-1, 2).copy()
        lib().orc_free_mem(rows)
        lib().orc_free_mem(succ)
        return r, s, json.loads(_take_str(actors))


class _Undefined:
    def __repr__(self):
        return 'Undefined'


Undefined = _Undefined()

KINDS = {'uint': 0, 'int': 1, 'utf8': 2, 'delta': 3, 'boolean': 4}


def decode_column(kind, data):
    out, err = C.c_void_p(), C.c_void_p()
    data = bytes(data)
    _check(lib().orc_decode_column(KINDS[kind], data, C.c_size_t(len(data)), C.byref(out), C.byref(err)), err.value)
    return json.loads(_take_str(out))


def encode_column(kind, values):
    n = len(values)
    vals = (C.c_int64 * max(n, 1))()
    nulls = (C.c_uint8 * max(n, 1))()
    strs, offs = b'', (C.c_uint64 * (n + 1))()
    for i, v in enumerate(values):
        offs[i] = len(strs)
        if v is None:
            nulls[i] = 1
        elif isinstance(v, str):
            strs += v.encode('utf-8')
        else:
            vals[i] = int(v)
    offs[n] = len(strs)
    out, ln, err = C.c_void_p(), C.c_size_t(), C.c_void_p()
    _check(lib().orc_encode_column(KINDS[kind], vals, nulls, strs, offs, C.c_size_t(n), C.byref(out), C.byref(ln), C.byref(err)), err.value)
    b = C.string_at(out, ln.value)
    lib().orc_free_mem(out)
    return b


def leb_decode(kind, data):
    """kind: 'uint53' | 'int53' | 'uint32' -> (value, consumed)"""
    k = {'uint53': 0, 'int53': 1, 'uint32': 2}[kind]
    v, n, err = C.c_int64(), C.c_size_t(), C.c_void_p()
    data = bytes(data)
    _check(lib().orc_leb_decode(k, data, C.c_size_t(len(data)), C.byref(v), C.byref(n), C.byref(err)), err.value)
    return v.value, n.value


def leb_encode(kind, value):
    k = {'uint53': 0, 'int53': 1}[kind]
    buf, n, err = (C.c_uint8 * 16)(), C.c_size_t(), C.c_void_p()
    _check(lib().orc_leb_encode(k, C.c_int64(value), buf, C.byref(n), C.byref(err)), err.value)
    return bytes(buf[:n.value])


def sha256(data):
    out = (C.c_uint8 * 32)()
    data = bytes(data)
    lib().orc_sha256(data, C.c_size_t(len(data)), out)
    return bytes(out)


def decode_change(data):
    out, err = C.c_void_p(), C.c_void_p()
    data = bytes(data)
    _check(lib().orc_decode_change(data, C.c_size_t(len(data)), C.byref(out), C.byref(err)), err.value)
    return json.loads(_take_str(out))
