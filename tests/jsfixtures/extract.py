#!/usr/bin/env python
"""Generates tests/golden/*.json from the reference's own mocha test files.

Runs /root/reference/test/{new_backend_test,backend_test,columnar_test}.js through the JS-subset
interpreter in jsmini.py with the reference's modules replaced by recording host objects:
every call the test makes on BackendDoc / Backend (applyChanges, getPatch, save, load, ...) and
every assertion is written out as a step with concrete arguments (binary changes as hex, expected
patches, expected column bytes, expected error patterns).  tests/replay.py re-executes those steps
against the CPU oracle (pinning it) and against the CUDA engine (parity), without needing
/root/reference or a JS engine at test time.

While extracting, the steps are executed against the oracle so that values the tests read back
(e.g. `backend.heads` used as deps of the next change) are available; any assertion the oracle
fails is reported at the end (non-zero exit) — that is the pinning signal during development.

usage: python tests/jsfixtures/extract.py [--ref /root/reference] [--out tests/golden]
"""
import argparse
import hashlib
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import jsmini  # noqa: E402
from jsmini import JSError, JSThrow, undefined  # noqa: E402
from automerge_classic_b200.columnar import encode_change, change_hash, decode_change, DecodeError  # noqa: E402
from automerge_classic_b200.backend import Backend, RangeError as FacadeRangeError  # noqa: E402
import oracle  # noqa: E402
from oracle import OracleDoc, OracleError  # noqa: E402

DOC_OPS_COLUMNS = {'objActor': 0x01, 'objCtr': 0x02, 'keyActor': 0x11, 'keyCtr': 0x13, 'keyStr': 0x15, 'idActor': 0x21,
                   'idCtr': 0x23, 'insert': 0x34, 'action': 0x42, 'valLen': 0x56, 'valRaw': 0x57, 'chldActor': 0x61,
                   'chldCtr': 0x63, 'succNum': 0x80, 'succActor': 0x81, 'succCtr': 0x83}


def to_json(v):
    v = jsmini.unwrap(v)
    if v is undefined or v is oracle.Undefined:
        return {'$undefined': True}
    if isinstance(v, (bytes, bytearray)):
        return {'$bytes': bytes(v).hex()}
    if isinstance(v, dict):
        return {k: to_json(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [to_json(x) for x in v]
    if isinstance(v, float) and v.is_integer() and abs(v) < 2 ** 53:
        return int(v)
    if isinstance(v, (HostDoc,)):
        return {'$doc': v.id}
    return v


def from_oracle(v):
    """oracle JSON patch -> python values ({'$bytes': hex} -> bytes)."""
    if isinstance(v, dict):
        if set(v.keys()) == {'$bytes'}:
            return bytes.fromhex(v['$bytes'])
        return {k: from_oracle(x) for k, x in v.items()}
    if isinstance(v, list):
        return [from_oracle(x) for x in v]
    return v


def deep_equal(a, b):
    a, b = jsmini.unwrap(a), jsmini.unwrap(b)
    if a is oracle.Undefined:
        a = undefined
    if b is oracle.Undefined:
        b = undefined
    if isinstance(a, dict) and isinstance(b, dict):
        ka = {k for k, v in a.items()}
        kb = {k for k, v in b.items()}
        return ka == kb and all(deep_equal(a[k], b[k]) for k in ka)
    if isinstance(a, (list, tuple)) and isinstance(b, (list, tuple)):
        return len(a) == len(b) and all(deep_equal(x, y) for x, y in zip(a, b))
    if isinstance(a, (bytes, bytearray)) and isinstance(b, (bytes, bytearray)):
        return bytes(a) == bytes(b)
    if isinstance(a, bool) or isinstance(b, bool):
        return a is b
    if isinstance(a, (int, float)) and isinstance(b, (int, float)):
        return a == b or (a != a and b != b)
    if a is None or b is None or a is undefined or b is undefined:
        return a is b
    return type(a) == type(b) and a == b


class Recorder:
    def __init__(self):
        self.tests, self.cur, self.failures, self.next_id = [], None, [], 0

    def new_id(self):
        self.next_id += 1
        return self.next_id

    def step(self, **kw):
        self.cur['steps'].append(kw)
        return kw


REC = Recorder()


class Traced:
    """A value that came out of the engine under test; remembers where it came from."""

    def __init__(self, res, path, value):
        self.res, self.path, self.value = res, path, value

    def js_unwrap(self):
        return self.value

    def ref(self):
        return {'res': self.res, 'path': list(self.path)}

    def js_get(self, key):
        v = self.value
        if isinstance(v, (list, bytes, bytearray, str)) and key == 'length':
            return Traced(self.res, self.path + ['length'], len(v))
        if isinstance(v, (bytes, bytearray)) and key == 'byteLength':
            return Traced(self.res, self.path + ['length'], len(v))
        if isinstance(v, dict):
            k = key if isinstance(key, str) else jsmini.jsstr(key)
            return Traced(self.res, self.path + [k], v.get(k, undefined))
        if isinstance(v, (list, bytes, bytearray)) and isinstance(key, (int, float)) and not isinstance(key, bool):
            i = int(key)
            return Traced(self.res, self.path + [i], v[i] if 0 <= i < len(v) else undefined)
        # fall back to plain semantics (methods such as .map / .slice): loses the trace
        return INTERP.getmember(v, key)

    def js_call(self, key, args):
        f = INTERP.getmember(self.value, key)
        if not callable(f):
            raise JSError('TypeError', '%s is not a function' % key)
        return f(*args)


def traced(value, **step):
    rid = REC.new_id()
    REC.step(res=rid, **step)
    return Traced(rid, [], value)


def ref_or_value(v):
    return {'ref': v.ref()} if isinstance(v, Traced) else {'value': to_json(v)}


def run_engine(step, fn):
    """Executes fn against the oracle; records the error on the step if it throws."""
    try:
        return fn()
    except OracleError as e:
        step['error'] = e.message
        raise JSError(e.kind, e.message)
    except FacadeRangeError as e:
        step['error'] = str(e)
        raise JSError('RangeError', str(e))
    except ValueError as e:
        step['error'] = str(e)
        raise JSError('RangeError', str(e))
    except (TypeError, RuntimeError) as e:
        step['error'] = str(e)
        raise JSError('TypeError' if isinstance(e, TypeError) else 'Error', str(e))


def as_bytes_list(changes):
    changes = jsmini.unwrap(changes)
    if isinstance(changes, (bytes, bytearray)):
        return changes
    if not isinstance(changes, list):
        return changes
    return [bytes(jsmini.unwrap(c)) for c in changes]


def hexlist(changes):
    if isinstance(changes, (bytes, bytearray)):
        return {'$bytes': bytes(changes).hex()}
    if not isinstance(changes, list):
        return to_json(changes)
    return [c.hex() for c in changes]


# ---------------------------------------------------------------- BackendDoc host object
class HostDoc:
    def __init__(self, doc, step_kw):
        self.id = REC.new_id()
        self.doc = doc
        REC.step(doc=self.id, **step_kw)

    def js_get(self, key):
        if key == 'blocks':
            return traced(self.doc.blocks(), op='blocks', doc=self.id)
        if key == 'heads':
            return traced(self.doc.heads(), op='heads', doc=self.id)
        if key == 'clock':
            return traced(self.doc.clock(), op='clock', doc=self.id)
        if key == 'maxOp':
            return traced(self.doc.max_op(), op='max_op', doc=self.id)
        if key in ('applyChanges', 'getPatch', 'save', 'clone', 'getChanges', 'getChangesAdded', 'getMissingDeps', 'getChangeByHash'):
            return lambda *a: self.js_call(key, list(a))
        raise JSError('Unsupported', 'BackendDoc.%s' % key)

    def js_call(self, key, args):
        if key == 'applyChanges':
            changes = as_bytes_list(args[0])
            local = bool(args[1]) if len(args) > 1 else False
            rid = REC.new_id()
            st = REC.step(op='apply', doc=self.id, changes=hexlist(changes), local=local, res=rid)
            if not isinstance(changes, list):
                st['error'] = 'applyChanges takes an array of Uint8Arrays'
                raise JSError('TypeError', st['error'] + (', not just a single Uint8Array' if isinstance(changes, (bytes, bytearray)) else ''))
            patch = run_engine(st, lambda: self.doc.apply_changes(changes, is_local=local))
            return Traced(rid, [], from_oracle(patch))
        if key == 'getPatch':
            rid = REC.new_id()
            st = REC.step(op='get_patch', doc=self.id, res=rid)
            return Traced(rid, [], from_oracle(run_engine(st, self.doc.get_patch)))
        if key == 'save':
            rid = REC.new_id()
            st = REC.step(op='save', doc=self.id, res=rid)
            return Traced(rid, [], run_engine(st, self.doc.save))
        if key == 'clone':
            return HostDoc(self.doc.clone(), dict(op='clone', src=self.id))
        if key == 'getChanges':
            rid = REC.new_id()
            deps = [jsmini.unwrap(x) for x in jsmini.unwrap(args[0])]
            st = REC.step(op='get_changes', doc=self.id, have_deps=deps, res=rid)
            return Traced(rid, [], run_engine(st, lambda: self.doc.get_changes(deps)))
        if key == 'getMissingDeps':
            rid = REC.new_id()
            heads = [jsmini.unwrap(x) for x in jsmini.unwrap(args[0])] if args else []
            st = REC.step(op='get_missing_deps', doc=self.id, heads=heads, res=rid)
            return Traced(rid, [], run_engine(st, lambda: self.doc.get_missing_deps(heads)))
        raise JSError('Unsupported', 'BackendDoc.%s()' % key)


def BackendDocCtor(buffer=None):
    buffer = jsmini.unwrap(buffer) if buffer is not None else None
    if buffer is None or buffer is undefined:
        return HostDoc(OracleDoc(), dict(op='new_doc'))
    data = bytes(buffer)
    st = {}
    try:
        doc = run_engine(st, lambda: OracleDoc(data))
    except JSError:
        REC.step(op='load_doc', doc=REC.new_id(), data=data.hex(), error=st.get('error'))
        raise
    return HostDoc(doc, dict(op='load_doc', data=data.hex()))


# ---------------------------------------------------------------- Backend facade host object
FACADE = Backend(OracleDoc)


class HostHandle:
    """A Backend handle ({state, heads, frozen}); wraps the facade's dict."""

    def __init__(self, h):
        self.id = REC.new_id()
        self.h = h

    def js_unwrap(self):
        return self


def _bstep(op, **kw):
    return REC.step(op=op, **kw)


def make_backend_module():
    def init():
        hh = HostHandle(FACADE.init())
        _bstep('b_init', h=hh.id)
        return hh

    def clone(b):
        hh = HostHandle(None)
        st = _bstep('b_clone', src=b.id, h=hh.id)
        hh.h = run_engine(st, lambda: FACADE.clone(b.h))
        return hh

    def free(b):
        _bstep('b_free', h=b.id)
        FACADE.free(b.h)
        return undefined

    def applyChanges(b, changes):
        changes = as_bytes_list(changes)
        h2, rid = HostHandle(None), REC.new_id()
        st = _bstep('b_apply', h=b.id, changes=hexlist(changes), h2=h2.id, res=rid)
        out = run_engine(st, lambda: FACADE.applyChanges(b.h, changes))
        h2.h = out[0]
        return [h2, Traced(rid, [], from_oracle(out[1]))]

    def applyLocalChange(b, change):
        change = to_plain(change)
        h2, rid, rid2 = HostHandle(None), REC.new_id(), REC.new_id()
        st = _bstep('b_apply_local', h=b.id, change=to_json(change), h2=h2.id, res=rid, res_bin=rid2)
        out = run_engine(st, lambda: FACADE.applyLocalChange(b.h, change))
        h2.h = out[0]
        return [h2, Traced(rid, [], from_oracle(out[1])), Traced(rid2, [], out[2])]

    def save(b):
        rid = REC.new_id()
        st = _bstep('b_save', h=b.id, res=rid)
        return Traced(rid, [], run_engine(st, lambda: FACADE.save(b.h)))

    def load(data):
        data = bytes(jsmini.unwrap(data))
        hh = HostHandle(None)
        st = _bstep('b_load', data=data.hex(), h=hh.id)
        hh.h = run_engine(st, lambda: FACADE.load(data))
        return hh

    def loadChanges(b, changes):
        changes = as_bytes_list(changes)
        h2 = HostHandle(None)
        st = _bstep('b_load_changes', h=b.id, changes=hexlist(changes), h2=h2.id)
        h2.h = run_engine(st, lambda: FACADE.loadChanges(b.h, changes))
        return h2

    def getPatch(b):
        rid = REC.new_id()
        st = _bstep('b_get_patch', h=b.id, res=rid)
        return Traced(rid, [], from_oracle(run_engine(st, lambda: FACADE.getPatch(b.h))))

    def getHeads(b):
        rid = REC.new_id()
        st = _bstep('b_get_heads', h=b.id, res=rid)
        return Traced(rid, [], run_engine(st, lambda: FACADE.getHeads(b.h)))

    def getAllChanges(b):
        rid = REC.new_id()
        st = _bstep('b_get_all_changes', h=b.id, res=rid)
        return Traced(rid, [], run_engine(st, lambda: FACADE.getAllChanges(b.h)))

    def getChanges(b, deps):
        deps = jsmini.unwrap(deps)
        deps = [jsmini.unwrap(x) for x in deps] if isinstance(deps, list) else deps
        rid = REC.new_id()
        st = _bstep('b_get_changes', h=b.id, have_deps=to_json(deps), res=rid)
        return Traced(rid, [], run_engine(st, lambda: FACADE.getChanges(b.h, deps)))

    def getChangesAdded(b1, b2):
        rid = REC.new_id()
        st = _bstep('b_get_changes_added', h1=b1.id, h2=b2.id, res=rid)
        return Traced(rid, [], run_engine(st, lambda: FACADE.getChangesAdded(b1.h, b2.h)))

    def getChangeByHash(b, h):
        rid = REC.new_id()
        h = jsmini.unwrap(h)
        st = _bstep('b_get_change_by_hash', h=b.id, hash=h, res=rid)
        v = run_engine(st, lambda: FACADE.getChangeByHash(b.h, h))
        return Traced(rid, [], undefined if v is None else v)

    def getMissingDeps(b, heads=None):
        rid = REC.new_id()
        heads = [jsmini.unwrap(x) for x in jsmini.unwrap(heads)] if heads is not None else []
        st = _bstep('b_get_missing_deps', h=b.id, heads=heads, res=rid)
        return Traced(rid, [], run_engine(st, lambda: FACADE.getMissingDeps(b.h, heads)))

    return {k: v for k, v in locals().items() if callable(v)}


def to_plain(v):
    v = jsmini.unwrap(v)
    if isinstance(v, dict):
        return {k: to_plain(x) for k, x in v.items()}
    if isinstance(v, list):
        return [to_plain(x) for x in v]
    if isinstance(v, bytearray):
        return bytes(v)
    return v


# ---------------------------------------------------------------- assertions & misc host functions
def js_encode_change(change):
    try:
        return encode_change(to_plain(change))
    except ValueError as e:
        raise JSError('RangeError', str(e))


def js_hash(change):
    return change_hash(to_plain(change))


def js_decode_change(buf, *_):
    try:
        return decode_change(bytes(jsmini.unwrap(buf)))
    except DecodeError as e:
        raise JSError('RangeError', str(e))


class JSDate:
    def __init__(self, ms=None):
        self.ms = 1600000000000 if ms is None else int(jsmini.unwrap(ms))

    def getTime(self):
        return self.ms


def record_assert(kind, actual, expected, ok):
    REC.step(op='assert_' + kind, actual=ref_or_value(actual), expected=ref_or_value(expected))
    if not ok:
        REC.failures.append((REC.cur['name'], kind, jsmini.unwrap(actual), jsmini.unwrap(expected)))


def assert_deep(actual, expected, msg=None):
    if not isinstance(actual, Traced) and not isinstance(expected, Traced):
        if not deep_equal(actual, expected):
            raise JSError('AssertionError', 'deepStrictEqual on plain values failed')
        return undefined
    record_assert('equal', actual, expected, deep_equal(actual, expected))
    return undefined


def assert_strict(actual, expected, msg=None):
    return assert_deep(actual, expected)


def assert_not_strict(actual, expected, msg=None):
    if isinstance(actual, Traced) or isinstance(expected, Traced):
        record_assert('not_equal', actual, expected, not deep_equal(actual, expected))
    return undefined


def assert_throws(fn, pattern=None, msg=None):
    n_before = len(REC.cur['steps'])
    try:
        fn()
    except (JSError, JSThrow) as e:
        text = e.message if isinstance(e, JSError) else jsmini.jsstr(getattr(e, 'value', ''))
        if isinstance(e, JSError) and e.kind == 'Unsupported':
            raise
        pat = pattern.pattern if isinstance(pattern, jsmini.JSRegex) else None
        steps = REC.cur['steps'][n_before:]
        target = next((s for s in reversed(steps) if 'error' in s), None)
        if target is None:
            # thrown by plain JS (not by the engine): nothing to replay
            return undefined
        target['throws'] = pat or ''
        if pat and not re.search(pat, text):
            REC.failures.append((REC.cur['name'], 'throws', text, pat))
        return undefined
    REC.failures.append((REC.cur['name'], 'throws', 'no exception', getattr(pattern, 'pattern', None)))
    return undefined


def assert_ok(v, msg=None):
    if isinstance(v, Traced):
        REC.step(op='assert_truthy', actual=v.ref())
        if not jsmini.truthy(v):
            REC.failures.append((REC.cur['name'], 'ok', v.value, True))
    elif not jsmini.truthy(v):
        raise JSError('AssertionError', 'assert(%r)' % (v,))
    return undefined


def check_columns(block, expected):
    expected = {k: bytes(bytearray(int(x) for x in jsmini.unwrap(v))) for k, v in jsmini.unwrap(expected).items()}
    REC.step(op='check_columns', block=block.ref(), expected={k: v.hex() for k, v in expected.items()})
    cols = block.value['columns']
    problems = []
    for cid, buf in cols.items():
        name = next((n for n, i in DOC_OPS_COLUMNS.items() if i == cid), str(cid))
        if name in expected:
            if buf != expected[name]:
                problems.append('%s: got %s expected %s' % (name, list(buf), list(expected[name])))
        elif name not in ('chldActor', 'chldCtr'):
            problems.append('unexpected column %s' % name)
    for name in expected:
        cid = DOC_OPS_COLUMNS.get(name, int(name) if name.isdigit() else -1)
        if cid not in cols:
            problems.append('missing column %s' % name)
    if problems:
        REC.failures.append((REC.cur['name'], 'checkColumns', problems, None))
    return undefined


def bloom_contains(bloom, actor, ctr):
    rid = REC.new_id()
    actor, ctr = int(jsmini.unwrap(actor)), int(jsmini.unwrap(ctr))
    REC.step(op='bloom_contains', bloom=bloom.ref(), actor=actor, ctr=ctr, res=rid)
    b = bloom.value
    modulo = 8 * len(b)
    x, y = ctr % modulo, actor % modulo
    z = (((ctr ^ actor) * 16777619) & 0xffffffff) % modulo
    ok = True
    for _ in range(7):
        if not b[x >> 3] & (1 << (x & 7)):
            ok = False
            break
        x = (x + y) % modulo
        y = (y + z) % modulo
    return Traced(rid, [], ok)


_uuid_counter = [0]


def uuid():
    _uuid_counter[0] += 1
    return hashlib.md5(b'amgpu-fixture-%d' % _uuid_counter[0]).hexdigest()


def uint8array(arg=None):
    arg = jsmini.unwrap(arg)
    if isinstance(arg, (int, float)):
        return bytearray(int(arg))
    return bytearray(int(jsmini.unwrap(x)) & 0xff for x in (arg or []))


def js_array(n=None):
    if n is None:
        return []
    return [undefined] * int(jsmini.unwrap(n))


class Unsupported:
    def __init__(self, name):
        self.name = name

    def js_get(self, key):
        if self.name == 'Automerge' and key == 'Backend':
            return BACKEND_MODULE
        return Unsupported(self.name + '.' + str(key))

    def js_call(self, key, args):
        raise JSError('Unsupported', '%s.%s()' % (self.name, key))

    def __call__(self, *a):
        raise JSError('Unsupported', self.name + '()')


BACKEND_MODULE = None
INTERP = None


def make_globals():
    global BACKEND_MODULE
    BACKEND_MODULE = make_backend_module()

    def describe(name, fn):
        STACK.append(name)
        try:
            fn()
        finally:
            STACK.pop()
        return undefined

    def it(name, fn):
        REC.cur = {'name': ' > '.join(STACK + [name]), 'steps': []}
        try:
            fn()
        except JSError as e:
            REC.cur['skipped'] = '%s: %s' % (e.kind, e.message)
        except Exception as e:   # interpreter limitation: keep going, report
            REC.cur['skipped'] = 'extractor error: %r' % (e,)
        REC.tests.append(REC.cur)
        return undefined

    def require(path):
        path = jsmini.unwrap(path)
        if path == 'assert':
            return ASSERT
        if path.endswith('helpers'):
            return {'checkEncoded': Unsupported('checkEncoded'), 'assertEqualsOneOf': Unsupported('assertEqualsOneOf')}
        if path.endswith('backend/columnar'):
            return {'DOC_OPS_COLUMNS': [], 'encodeChange': js_encode_change, 'decodeChange': js_decode_change,
                    'decodeChanges': Unsupported('decodeChanges')}
        if path.endswith('backend/new'):
            return {'MAX_BLOCK_SIZE': 600, 'BackendDoc': BackendDocCtor, 'bloomFilterContains': bloom_contains}
        if path.endswith('src/uuid'):
            return uuid
        if path.endswith('automerge'):
            return Unsupported('Automerge')
        return Unsupported(path)

    class AssertObj:
        def js_get(self, key):
            return {'deepStrictEqual': assert_deep, 'strictEqual': assert_strict, 'notStrictEqual': assert_not_strict,
                    'throws': assert_throws, 'ok': assert_ok, 'deepEqual': assert_deep, 'equal': assert_strict}[key]

        def js_call(self, key, args):
            return self.js_get(key)(*args)

        def __call__(self, *a):
            return assert_ok(*a)
    ASSERT = AssertObj()
    STACK = []

    def object_assign(target, *srcs):
        for s in srcs:
            s = jsmini.unwrap(s)
            if isinstance(s, dict):
                target.update(s)
        return target

    g = {
        'require': require, 'describe': describe, 'it': it, 'assert': ASSERT,
        'process': {'env': {}}, 'Uint8Array': uint8array, 'Array': js_array,
        'Object': {'keys': lambda o: list(jsmini.unwrap(o).keys()), 'assign': object_assign,
                   'entries': lambda o: [[k, v] for k, v in jsmini.unwrap(o).items()],
                   'values': lambda o: list(jsmini.unwrap(o).values())},
        'Math': {'floor': lambda x: int(jsmini.tonum(x) // 1), 'ceil': lambda x: -int(-jsmini.tonum(x) // 1),
                 'min': lambda *a: min(jsmini.tonum(x) for x in a), 'max': lambda *a: max(jsmini.tonum(x) for x in a),
                 'pow': lambda a, b: jsmini.norm(jsmini.tonum(a) ** jsmini.tonum(b)), 'round': lambda x: int(jsmini.tonum(x) + 0.5)},
        'parseInt': lambda s, base=10: int(jsmini.jsstr(s), int(base)),
        'Number': {'MAX_SAFE_INTEGER': 2 ** 53 - 1, 'MIN_SAFE_INTEGER': -(2 ** 53 - 1)},
        'JSON': {'stringify': lambda v: json.dumps(to_json(v))},
        'checkColumns': check_columns, 'hash': js_hash,
        'console': {'log': lambda *a: undefined}, 'Date': JSDate,
    }
    return g


def run_file(path):
    global INTERP
    REC.tests = []
    src = open(path).read()
    INTERP = jsmini.Interp(make_globals())
    INTERP.protected = {'checkColumns', 'hash'}
    INTERP.run(src)
    return REC.tests


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ref', default='/root/reference')
    ap.add_argument('--out', default=os.path.join(ROOT, 'tests', 'golden'))
    ap.add_argument('--files', nargs='*', default=['new_backend_test.js', 'backend_test.js'])
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    for f in args.files:
        tests = run_file(os.path.join(args.ref, 'test', f))
        skipped = [t for t in tests if 'skipped' in t]
        out = os.path.join(args.out, f.replace('.js', '.json'))
        with open(out, 'w') as fh:
            json.dump({'source': 'test/' + f, 'generator': 'tests/jsfixtures/extract.py', 'tests': tests}, fh, separators=(',', ':'))
        print('%s: %d tests, %d skipped -> %s' % (f, len(tests), len(skipped), out))
        for t in skipped:
            print('   skipped: %s  [%s]' % (t['name'], t['skipped'][:120]))
    if REC.failures:
        print('\n%d oracle mismatches:' % len(REC.failures))
        for name, kind, a, b in REC.failures[:40]:
            print(' -', name, '|', kind)
            print('     actual  :', json.dumps(to_json(a))[:1500])
            print('     expected:', json.dumps(to_json(b))[:1500])
        sys.exit(1)


if __name__ == '__main__':
    main()
