"""ctypes binding of libamgpu.so (include/amgpu.h) and the document class the Backend facade drives.

`GpuBackendDoc` has the surface of the reference's BackendDoc (backend/new.js:1694-2069): the heavy
lifting — columnar decode, causal gate, op-set ordering, patch computation — runs in the CUDA kernels
behind the C ABI; this module only (1) passes byte buffers across the boundary and (2) inflates the
flat binary patch table into the nested Patch object of @types/automerge/index.d.ts:236-316, which
is what the N-API shim does on the JavaScript side.

There is no CPU fallback: importing works anywhere, but creating a document raises if libamgpu.so is
missing or no CUDA device is present.
"""
import ctypes as C
import os
import struct

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libamgpu.so')

ACTIONS = ['makeMap', 'set', 'makeList', 'del', 'makeText', 'inc', 'makeTable', 'link']
OBJ_TYPE = {0: 'map', 2: 'list', 4: 'text', 6: 'table'}
PROP_DT = np.dtype([('obj', '<u8'), ('opId', '<u8'), ('keyOff', '<u4'), ('keyLen', '<u4'), ('valLen', '<u4'), ('valOff', '<u4'), ('flags', '<u4'), ('pad', '<u4')])
EDIT_DT = np.dtype([('obj', '<u8'), ('opId', '<u8'), ('index', '<u4'), ('kind', '<u4'), ('valLen', '<u4'), ('valOff', '<u4')])


class AmgError(Exception):
    """Error raised by the engine; `.kind` is the reference's JS error class."""
    KINDS = {1: 'RangeError', 2: 'TypeError', 3: 'Error', 4: 'Unsupported', 5: 'CudaError'}

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code, self.kind, self.message = code, self.KINDS.get(code, 'Error'), msg


class Unsupported(AmgError):
    pass


class _ErrStruct(C.Structure):
    _fields_ = [('code', C.c_int), ('msg', C.c_char * 512)]


class Library:
    def __init__(self, path):
        if not os.path.exists(path):
            raise RuntimeError('amgpu: %s not found — build it with `python -c "import __graft_entry__ as g; g.build()"` '
                               '(there is no CPU fallback)' % path)
        self.path = path
        L = self.L = C.CDLL(path)
        vp = C.c_void_p
        L.amg_init.restype = vp
        L.amg_init.argtypes = [C.c_int, vp]
        L.amg_load.restype = vp
        L.amg_load.argtypes = [C.c_int, vp, C.c_size_t, vp]
        L.amg_clone.restype = vp
        L.amg_clone.argtypes = [vp, vp]
        L.amg_free.argtypes = [vp]
        L.amg_patch_bytes.restype = vp
        L.amg_patch_bytes.argtypes = [vp, vp]
        L.amg_patch_free.argtypes = [vp]
        L.amg_arena.restype = vp
        L.amg_arena.argtypes = [vp, vp]
        L.amg_buffers_count.restype = C.c_size_t
        L.amg_buffers_count.argtypes = [vp]
        L.amg_buffers_get.restype = vp
        L.amg_buffers_get.argtypes = [vp, C.c_size_t, vp]
        L.amg_buffers_free.argtypes = [vp]
        L.amg_kernel_launches.restype = C.c_uint64
        L.amg_kernel_launches.argtypes = [vp]
        L.amg_free_mem.argtypes = [vp]
        for name in ('amg_apply_changes', 'amg_apply_changes_packed', 'amg_get_patch', 'amg_get_state', 'amg_get_heads', 'amg_get_changes',
                     'amg_get_changes_added', 'amg_get_change_by_hash', 'amg_get_missing_deps', 'amg_clock_of', 'amg_hash_by_actor',
                     'amg_debug_dump_ops', 'amg_debug_decode', 'amg_debug_decode_column', 'amg_bench_decode', 'amg_last_timings'):
            getattr(L, name).restype = C.c_int

    def check(self, rc, err):
        if rc != 0:
            msg = err.msg.decode('utf-8', 'replace')
            raise (Unsupported if rc == 4 else AmgError)(rc, msg)


_default = None


def default_library():
    global _default
    if _default is None:
        _default = Library(LIB_PATH)
    return _default


def decode_value(val_len, raw):
    """reference columnar.js:300-329 decodeValue -> patch value dict {'type':'value','value':..,['datatype':..]}"""
    tag = val_len & 15
    if val_len == 0:
        return {'type': 'value', 'value': None}
    if val_len == 1:
        return {'type': 'value', 'value': False}
    if val_len == 2:
        return {'type': 'value', 'value': True}
    if tag == 6:
        return {'type': 'value', 'value': bytes(raw).decode('utf-8', 'replace')}
    if tag in (3, 4, 8, 9):
        v, shift = 0, 0
        for b in raw:
            v |= (b & 0x7f) << shift
            shift += 7
            if not b & 0x80:
                if tag != 3 and b & 0x40:
                    v -= 1 << shift
                break
        return {'type': 'value', 'value': v, 'datatype': {3: 'uint', 4: 'int', 8: 'counter', 9: 'timestamp'}[tag]}
    if tag == 5:
        if len(raw) != 8:
            raise AmgError(1, 'Invalid length for floating point number: %d' % len(raw))
        return {'type': 'value', 'value': struct.unpack('<d', bytes(raw))[0], 'datatype': 'float64'}
    return {'type': 'value', 'value': bytes(raw), 'datatype': tag}


def empty_object_patch(object_id, action):
    t = OBJ_TYPE.get(action)
    if t in ('list', 'text'):
        return {'objectId': object_id, 'type': t, 'edits': []}
    return {'objectId': object_id, 'type': t, 'props': {}}


class FlatPatch:
    """Zero-copy view of the flat patch table (layout in include/amgpu.h)."""

    def __init__(self, raw):
        self.raw = self.arena = raw   # keyOff / valOff of the records index the patch buffer itself (its bytes section)
        h = self.hdr = np.frombuffer(raw, dtype='<u8', count=20)
        assert int(h[0]) == 0x31504747414d41, 'bad patch magic'
        self.max_op, self.pending = int(h[1]), int(h[2])
        self.actor_seq = None
        if int(h[3]):
            self.actor_seq = (bytes(raw[int(h[5]):int(h[5]) + int(h[6])]).hex(), int(h[4]))
        self.actors, off = [], int(h[7])
        for _ in range(int(h[8])):
            ln = struct.unpack_from('<I', raw, off)[0]
            self.actors.append(bytes(raw[off + 4:off + 4 + ln]).hex())
            off += 4 + ln
            off += (-off) % 4
        ck = np.frombuffer(raw, dtype='<u8', count=2 * int(h[10]), offset=int(h[9])).reshape(-1, 2)
        self.clock = {self.actors[int(a)]: int(s) for a, s in ck}
        self.deps = [bytes(raw[int(h[11]) + 32 * i:int(h[11]) + 32 * i + 32]).hex() for i in range(int(h[12]))]
        self.props = np.frombuffer(raw, dtype=PROP_DT, count=int(h[14]), offset=int(h[13]))
        self.edits = np.frombuffer(raw, dtype=EDIT_DT, count=int(h[16]), offset=int(h[15]))
        # elemOff == 0: the section is not shipped because every insert's elemId equals its opId
        self.edit_elem = np.frombuffer(raw, dtype='<u8', count=int(h[16]), offset=int(h[17])) if int(h[17]) else self.edits['opId']

    def op_id(self, x):
        x = int(x)
        return '%d@%s' % (x >> 16, self.actors[x & 0xffff])

    def header(self):
        out = {'maxOp': self.max_op, 'clock': self.clock, 'deps': self.deps, 'pendingChanges': self.pending}
        if self.actor_seq:
            out['actor'], out['seq'] = self.actor_seq
        return out

    def to_patch(self, whole_doc):
        """Assembles the nested Patch (what the N-API shim does in JS)."""
        arena = self.arena
        patches = {'_root': {'objectId': '_root', 'type': 'map', 'props': {}}}
        # pass 1: every object that appears as a value gets its (empty) patch
        for rec in self.props:
            action = int(rec['flags']) >> 8
            if action % 2 == 0 and not int(rec['flags']) & 1:
                oid = self.op_id(rec['opId'])
                patches.setdefault(oid, empty_object_patch(oid, action))
        for rec in self.edits:
            action = int(rec['kind']) >> 16
            if action % 2 == 0 and (int(rec['kind']) & 0xff) != 1:
                oid = self.op_id(rec['opId'])
                patches.setdefault(oid, empty_object_patch(oid, action))
        # pass 2: map entries
        for rec in self.props:
            obj = '_root' if int(rec['obj']) == 0 else self.op_id(rec['obj'])
            p = patches.get(obj)
            if p is None or 'props' not in p:
                continue   # object not reachable from the root (its make op is no longer visible)
            key = bytes(arena[int(rec['keyOff']):int(rec['keyOff']) + int(rec['keyLen'])]).decode('utf-8', 'replace')
            flags = int(rec['flags'])
            action = flags >> 8
            if flags & 1:
                p['props'].setdefault(key, {})
            elif action == 1 and flags & 2:     # counter: the engine summed the increments (new.js:941-966)
                total = int(rec['valOff']) | (int(rec['pad']) << 32)
                total -= (1 << 64) if total >= (1 << 63) else 0
                p['props'].setdefault(key, {})[self.op_id(rec['opId'])] = {'type': 'value', 'value': total, 'datatype': 'counter'}
            elif action == 1:
                vl, vo = int(rec['valLen']), int(rec['valOff'])
                p['props'].setdefault(key, {})[self.op_id(rec['opId'])] = decode_value(vl, arena[vo:vo + (vl >> 4)])
            elif action % 2 == 0:
                oid = self.op_id(rec['opId'])
                p['props'].setdefault(key, {})[oid] = patches[oid]
            elif not whole_doc:
                p['props'].setdefault(key, {})
        # pass 3: list edits (already ordered per object; runs flagged by the RunFlag kernel)
        for j, rec in enumerate(self.edits):
            obj = self.op_id(rec['obj'])
            p = patches.get(obj)
            if p is None or 'edits' not in p:
                continue
            edits = p['edits']
            kind, run_start, multi, action = int(rec['kind']) & 0xff, bool(int(rec['kind']) & 0x100), bool(int(rec['kind']) & 0x200), int(rec['kind']) >> 16
            index = int(rec['index'])
            if kind == 1:
                if run_start:
                    edits.append({'action': 'remove', 'index': index, 'count': 1})
                else:
                    edits[-1]['count'] += 1
                continue
            if action == 1 and int(rec['kind']) & 0x1000:   # counter with increments: the engine summed them (new.js:941-966)
                total = int(rec['valLen']) | (int(rec['valOff']) << 32)
                total -= (1 << 64) if total >= (1 << 63) else 0
                value = {'type': 'value', 'value': total, 'datatype': 'counter'}
            elif action == 1:
                vl, vo = int(rec['valLen']), int(rec['valOff'])
                value = decode_value(vl, arena[vo:vo + (vl >> 4)])
            elif action % 2 == 0:
                value = patches[self.op_id(rec['opId'])]
            else:
                continue
            if kind == 2:
                edits.append({'action': 'update', 'index': index, 'opId': self.op_id(rec['opId']), 'value': value})
            elif not run_start:
                edits[-1]['values'].append(value['value'])     # continues the multi-insert opened by an earlier record
            elif multi:
                edit = {'action': 'multi-insert', 'index': index, 'elemId': self.op_id(self.edit_elem[j]), 'values': [value['value']]}
                if value.get('datatype'):
                    edit['datatype'] = value['datatype']
                edits.append(edit)
            else:
                edits.append({'action': 'insert', 'index': index, 'elemId': self.op_id(self.edit_elem[j]), 'opId': self.op_id(rec['opId']), 'value': value})
        out = self.header()
        out['diffs'] = patches['_root']
        return out


class GpuBackendDoc:
    """BackendDoc (backend/new.js:1694) over the CUDA engine."""
    _library = None   # tests may bind a subclass to another build of the same sources

    @classmethod
    def lib(cls):
        return cls._library or default_library()

    def __init__(self, data=None, _handle=None, device=0):
        self._lib = self.lib()
        L = self._lib.L
        if _handle is not None:
            self.h = _handle
            return
        err = _ErrStruct()
        dev = int(os.environ.get('AMG_DEVICE', device))
        if data is not None:
            data = bytes(data)
            h = L.amg_load(dev, data, C.c_size_t(len(data)), C.byref(err))
        else:
            h = L.amg_init(dev, C.byref(err))
        if not h:
            raise (Unsupported if err.code == 4 else AmgError)(err.code, err.msg.decode('utf-8', 'replace'))
        self.h = C.c_void_p(h)

    def __del__(self):
        try:
            if getattr(self, 'h', None):
                self._lib.L.amg_free(self.h)
                self.h = None
        except Exception:
            pass

    # ---- helpers
    def _arena(self):
        n = C.c_size_t()
        p = self._lib.L.amg_arena(self.h, C.byref(n))
        if not p or n.value == 0:
            return memoryview(b'')
        return memoryview((C.c_uint8 * n.value).from_address(p)).cast('B')

    def _take_patch(self, pp):
        n = C.c_size_t()
        p = self._lib.L.amg_patch_bytes(pp, C.byref(n))
        raw = bytes((C.c_uint8 * n.value).from_address(p))
        self._lib.L.amg_patch_free(pp)
        return FlatPatch(raw)

    def _buffers(self, bl):
        L = self._lib.L
        out = []
        for i in range(L.amg_buffers_count(bl)):
            n = C.c_size_t()
            p = L.amg_buffers_get(bl, i, C.byref(n))
            out.append(bytes((C.c_uint8 * n.value).from_address(p)) if n.value else b'')
        L.amg_buffers_free(bl)
        return out

    # ---- BackendDoc surface
    def clone(self):
        err = _ErrStruct()
        h = self._lib.L.amg_clone(self.h, C.byref(err))
        if not h:
            raise AmgError(err.code, err.msg.decode('utf-8', 'replace'))
        return type(self)(_handle=C.c_void_p(h))

    def apply_changes_flat(self, changes, is_local=False, want_patch=True):
        n = len(changes)
        blob = b''.join(bytes(c) for c in changes)
        offs = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum([len(c) for c in changes], out=offs[1:])
        return self.apply_packed_flat(blob, offs, n, is_local, want_patch)

    def apply_packed_flat(self, blob, offs, n, is_local=False, want_patch=True):
        pp, err = C.c_void_p(), _ErrStruct()
        if isinstance(blob, (bytes, bytearray)):
            buf = (C.c_uint8 * max(len(blob), 1)).from_buffer_copy(bytes(blob) if len(blob) else b'\0')   # (an empty list of changes is legal)
        elif isinstance(blob, np.ndarray):
            buf = blob.ctypes.data_as(C.c_void_p)
        else:
            buf = blob   # a ctypes pointer / address (e.g. pinned host memory)
        rc = self._lib.L.amg_apply_changes_packed(self.h, buf, offs.ctypes.data_as(C.c_void_p), C.c_size_t(n), int(is_local), int(want_patch),
                                                  C.byref(pp), C.byref(err))
        self._lib.check(rc, err)
        return self._take_patch(pp) if want_patch else None

    def apply_changes_ptrs_flat(self, changes, is_local=False, want_patch=True):
        """amg_apply_changes: one pointer + length per change (what an N-API binding passes for a Uint8Array[])."""
        n = len(changes)
        keep = [bytes(c) for c in changes]
        bufs = (C.c_char_p * max(n, 1))(*keep)
        lens = (C.c_size_t * max(n, 1))(*[len(c) for c in keep])
        pp, err = C.c_void_p(), _ErrStruct()
        fn = self._lib.L.amg_apply_changes
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        rc = fn(self.h, bufs, lens, C.c_size_t(n), int(is_local), int(want_patch), C.byref(pp), C.byref(err))
        self._lib.check(rc, err)
        return self._take_patch(pp) if want_patch else None

    def apply_changes(self, changes, is_local=False, want_patch=True):
        if isinstance(changes, (bytes, bytearray)):
            raise TypeError('applyChanges takes an array of Uint8Arrays, not just a single Uint8Array')
        fp = self.apply_changes_flat(list(changes), is_local, want_patch)
        return fp.to_patch(False) if want_patch else None

    def get_patch_flat(self):
        pp, err = C.c_void_p(), _ErrStruct()
        self._lib.check(self._lib.L.amg_get_patch(self.h, C.byref(pp), C.byref(err)), err)
        return self._take_patch(pp)

    def get_patch(self):
        return self.get_patch_flat().to_patch(True)

    def _state(self):
        pp, err = C.c_void_p(), _ErrStruct()
        self._lib.check(self._lib.L.amg_get_state(self.h, C.byref(pp), C.byref(err)), err)
        return self._take_patch(pp)

    def heads(self):
        return self._state().deps

    def clock(self):
        return self._state().clock

    def max_op(self):
        return self._state().max_op

    def save(self):
        """Backend.save (new.js:2033-2055): the document chunk, columns encoded on the device."""
        fn = getattr(self._lib.L, 'amg_save', None)
        if fn is None:
            raise Unsupported(4, 'amgpu: this build of libamgpu has no amg_save')
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        bl, err = C.c_void_p(), _ErrStruct()
        self._lib.check(fn(self.h, C.byref(bl), C.byref(err)), err)
        return self._buffers(bl)[0]

    def hash_by_actor(self, actor, index):
        a = bytes.fromhex(actor)
        out, found, err = (C.c_uint8 * 32)(), C.c_int(), _ErrStruct()
        self._lib.check(self._lib.L.amg_hash_by_actor(self.h, a, C.c_size_t(len(a)), C.c_uint64(index), out, C.byref(found), C.byref(err)), err)
        return bytes(out).hex() if found.value else None

    def get_changes(self, have_deps):
        deps = b''.join(bytes.fromhex(h) for h in have_deps)
        bl, err = C.c_void_p(), _ErrStruct()
        self._lib.check(self._lib.L.amg_get_changes(self.h, deps, C.c_size_t(len(have_deps)), C.byref(bl), C.byref(err)), err)
        return self._buffers(bl)

    def get_changes_added(self, old):
        bl, err = C.c_void_p(), _ErrStruct()
        self._lib.check(self._lib.L.amg_get_changes_added(self.h, old.h, C.byref(bl), C.byref(err)), err)
        return self._buffers(bl)

    def get_change_by_hash(self, hash_):
        bl, err = C.c_void_p(), _ErrStruct()
        self._lib.check(self._lib.L.amg_get_change_by_hash(self.h, bytes.fromhex(hash_), C.byref(bl), C.byref(err)), err)
        got = self._buffers(bl)
        return got[0] if got else None

    def get_missing_deps(self, heads=()):
        hs = b''.join(bytes.fromhex(h) for h in heads)
        bl, err = C.c_void_p(), _ErrStruct()
        self._lib.check(self._lib.L.amg_get_missing_deps(self.h, hs, C.c_size_t(len(heads)), C.byref(bl), C.byref(err)), err)
        return [b.hex() for b in self._buffers(bl)]

    def dump_ops(self):
        rows, n, succ, m, err = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t(), _ErrStruct()
        self._lib.check(self._lib.L.amg_debug_dump_ops(self.h, C.byref(rows), C.byref(n), C.byref(succ), C.byref(m), C.byref(err)), err)
        r = np.frombuffer(C.string_at(rows, 8 * 8 * n.value), dtype=np.uint64).reshape(-1, 8).copy()
        s = np.frombuffer(C.string_at(succ, 8 * 2 * m.value), dtype=np.uint64).reshape(-1, 2).copy()
        self._lib.L.amg_free_mem(rows)
        self._lib.L.amg_free_mem(succ)
        return r, s

    def debug_decode(self, changes):
        """Decoded rows of a batch of binary changes straight from the decode kernels (amg_debug_decode): the document is
        not touched. Returns (hashes [n] bytes, n_ops [n], rows {column: uint32 array in batch order}, preds {'predActor',
        'predCtr'}, staged bytes that keyStrOff / valOff index: the changes back to back, DEFLATEd ones inflated)."""
        n = len(changes)
        blob = b''.join(bytes(c) for c in changes)
        offs = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum([len(c) for c in changes], out=offs[1:])
        buf = (C.c_uint8 * max(len(blob), 1)).from_buffer_copy(blob if blob else b'\0')
        hashes = (C.c_uint8 * (32 * max(n, 1)))()
        n_ops = np.zeros(max(n, 1), dtype=np.uint32)
        rows, total, total_preds, err = C.c_void_p(), C.c_size_t(), C.c_size_t(), _ErrStruct()
        self._lib.check(self._lib.L.amg_debug_decode(self.h, buf, offs.ctypes.data_as(C.c_void_p), C.c_size_t(n), hashes, n_ops.ctypes.data_as(C.c_void_p),
                                                     C.byref(rows), C.byref(total), C.byref(total_preds), C.byref(err)), err)
        M, P = total.value, total_preds.value
        flat = np.frombuffer(C.string_at(rows, 4 * (12 * M + 2 * P)), dtype=np.uint32).copy()
        self._lib.L.amg_free_mem(rows)
        names = ['objActor', 'objCtr', 'keyActor', 'keyCtr', 'keyStrOff', 'keyStrLen', 'insert', 'action', 'valLen', 'valOff', 'predNum', 'predOff']
        cols = {name: flat[k * M:(k + 1) * M] for k, name in enumerate(names)}
        preds = {'predActor': flat[12 * M:12 * M + P], 'predCtr': flat[12 * M + P:12 * M + 2 * P]}
        hs = bytes(hashes)
        return [hs[32 * i:32 * i + 32] for i in range(n)], n_ops[:n].copy(), cols, preds

    def debug_decode_column(self, buf, kind, n, parallel):
        """One document column through the parallel (True) or serial decoder; (rc, values, message), rc 1 = declined."""
        out = np.zeros(max(n, 1), dtype=np.int64)
        err = _ErrStruct()
        raw = bytes(buf)
        b = (C.c_uint8 * max(len(raw), 1)).from_buffer_copy(raw if raw else b'\0')
        rc = self._lib.L.amg_debug_decode_column(self.h, b, C.c_size_t(len(raw)), C.c_int(kind), C.c_size_t(n), C.c_int(1 if parallel else 0),
                                                 out.ctypes.data_as(C.c_void_p), C.byref(err))
        return rc, out[:n].tolist(), err.msg.decode('utf-8', 'replace')

    def timings(self):
        out = (C.c_float * 24)()
        self._lib.L.amg_last_timings(self.h, out, 24)
        return list(out)

    def launches(self):
        return int(self._lib.L.amg_kernel_launches(self.h))


def doc_class_for(library_path):
    """A GpuBackendDoc subclass bound to another build of libamgpu (tests only)."""
    lib = Library(library_path)
    return type('BoundBackendDoc', (GpuBackendDoc,), {'_library': lib})
