"""Replays the golden fixtures extracted from the reference's mocha tests (tests/golden/*.json, made
by tests/jsfixtures/extract.py) against a document engine.

`doc_class` is any class with the BackendDoc surface (oracle.OracleDoc, or the CUDA engine's
GpuBackendDoc).  Steps that inspect the reference's internal block structure (`blocks`,
`check_columns`, `bloom_contains`) only make sense for the oracle — `structural=False` skips them.
"""
import json
import os
import re

from automerge_classic_b200.backend import Backend

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
DOC_OPS_COLUMNS = {'objActor': 0x01, 'objCtr': 0x02, 'keyActor': 0x11, 'keyCtr': 0x13, 'keyStr': 0x15, 'idActor': 0x21,
                   'idCtr': 0x23, 'insert': 0x34, 'action': 0x42, 'valLen': 0x56, 'valRaw': 0x57, 'chldActor': 0x61,
                   'chldCtr': 0x63, 'succNum': 0x80, 'succActor': 0x81, 'succCtr': 0x83}


class Undef:
    def __repr__(self):
        return 'undefined'


UNDEF = Undef()


def load(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)['tests']


def decode(v):
    """fixture JSON -> python ({'$bytes'}: bytes, {'$undefined'}: UNDEF)."""
    if isinstance(v, dict):
        if set(v.keys()) == {'$bytes'}:
            return bytes.fromhex(v['$bytes'])
        if set(v.keys()) == {'$undefined'}:
            return UNDEF
        return {k: decode(x) for k, x in v.items()}
    if isinstance(v, list):
        return [decode(x) for x in v]
    return v


def deep_equal(a, b, path=''):
    """Returns None if equal (assert.deepStrictEqual semantics: key order irrelevant, arrays ordered),
    else a string describing the first difference."""
    if type(a).__name__ in ('_Undefined', 'Undef'):
        a = UNDEF
    if type(b).__name__ in ('_Undefined', 'Undef'):
        b = UNDEF
    if isinstance(a, dict) and isinstance(b, dict):
        if set(a.keys()) != set(b.keys()):
            return '%s: keys %s != %s' % (path, sorted(a.keys()), sorted(b.keys()))
        for k in a:
            d = deep_equal(a[k], b[k], path + '.' + k)
            if d:
                return d
        return None
    if isinstance(a, (list, tuple)) and isinstance(b, (list, tuple)):
        if len(a) != len(b):
            return '%s: length %d != %d\n   actual   %r\n   expected %r' % (path, len(a), len(b), a[:6], b[:6])
        for i, (x, y) in enumerate(zip(a, b)):
            d = deep_equal(x, y, '%s[%d]' % (path, i))
            if d:
                return d
        return None
    if isinstance(a, (bytes, bytearray)) and isinstance(b, (bytes, bytearray)):
        return None if bytes(a) == bytes(b) else '%s: bytes %s != %s' % (path, bytes(a).hex(), bytes(b).hex())
    if isinstance(a, bool) or isinstance(b, bool):
        return None if a is b else '%s: %r != %r' % (path, a, b)
    if isinstance(a, (int, float)) and isinstance(b, (int, float)):
        return None if (a == b or (a != a and b != b)) else '%s: %r != %r' % (path, a, b)
    if a is None or b is None or a is UNDEF or b is UNDEF:
        return None if a is b else '%s: %r != %r' % (path, a, b)
    if type(a) != type(b) or a != b:
        return '%s: %r != %r' % (path, a, b)
    return None


def bloom_contains(b, actor, ctr):
    modulo = 8 * len(b)
    x, y = ctr % modulo, actor % modulo
    z = (((ctr ^ actor) * 16777619) & 0xffffffff) % modulo
    for _ in range(7):
        if not b[x >> 3] & (1 << (x & 7)):
            return False
        x = (x + y) % modulo
        y = (y + z) % modulo
    return True


class Replayer:
    def __init__(self, doc_class, engine_errors, structural=True):
        self.doc_class, self.engine_errors, self.structural = doc_class, engine_errors, structural
        self.facade = Backend(doc_class)

    def resolve(self, ref):
        v = self.results[ref['res']]
        for p in ref['path']:
            if p == 'length' and isinstance(v, (list, bytes, bytearray, str)):
                v = len(v)
            elif isinstance(v, dict):
                v = v.get(p, UNDEF)
            else:
                v = v[p] if 0 <= p < len(v) else UNDEF
        return v

    def side(self, x):
        return self.resolve(x['ref']) if 'ref' in x else decode(x['value'])

    def run_test(self, test):
        """Returns a list of failure strings (empty = pass)."""
        self.results, self.docs, self.handles, fails = {}, {}, {}, []
        skip_refs = set()

        def guarded(step, fn):
            try:
                out = fn()
            except self.engine_errors as e:
                msg = str(e)
                if 'throws' in step or 'error' in step:
                    pat = step.get('throws')
                    if pat and not re.search(pat, msg):
                        fails.append('step %s: error %r does not match /%s/' % (step['op'], msg, pat))
                    return None, True
                fails.append('step %s: unexpected error %s' % (step['op'], msg))
                return None, True
            if 'throws' in step or 'error' in step:
                fails.append('step %s: expected error /%s/ but call succeeded' % (step['op'], step.get('throws', step.get('error'))))
            return out, False

        def changes_of(step):
            c = step['changes']
            if isinstance(c, dict):
                return decode(c)
            return [bytes.fromhex(x) for x in c]

        for step in test['steps']:
            op = step['op']
            # a step whose inputs were never produced (an earlier step failed) cannot run
            needs = [step.get(k) for k in ('h', 'src', 'h1') if k in step and op.startswith('b_') and op not in ('b_init', 'b_load')]
            if op == 'b_get_changes_added':
                needs.append(step['h2'])
            if any(x not in self.handles or self.handles[x] is None for x in needs):
                if not fails:
                    fails.append('step %s: input handle missing' % op)
                continue
            if op in ('apply', 'get_patch', 'save', 'heads', 'clock', 'max_op', 'get_changes', 'get_missing_deps', 'blocks', 'clone') and \
                    (step.get('doc', step.get('src')) not in self.docs):
                if not fails:
                    fails.append('step %s: input doc missing' % op)
                continue
            if op == 'new_doc':
                self.docs[step['doc']] = self.doc_class()
            elif op == 'load_doc':
                out, failed = guarded(step, lambda: self.doc_class(bytes.fromhex(step['data'])))
                if not failed:
                    self.docs[step['doc']] = out
            elif op == 'clone':
                self.docs[step['doc']] = self.docs[step['src']].clone()
            elif op == 'apply':
                out, failed = guarded(step, lambda: self.docs[step['doc']].apply_changes(changes_of(step), is_local=step['local']))
                self.results[step['res']] = decode(out) if not failed else UNDEF
            elif op == 'get_patch':
                self.results[step['res']] = decode(self.docs[step['doc']].get_patch())
            elif op == 'save':
                self.results[step['res']] = self.docs[step['doc']].save()
            elif op == 'heads':
                self.results[step['res']] = self.docs[step['doc']].heads()
            elif op == 'clock':
                self.results[step['res']] = self.docs[step['doc']].clock()
            elif op == 'max_op':
                self.results[step['res']] = self.docs[step['doc']].max_op()
            elif op == 'get_changes':
                out, failed = guarded(step, lambda: self.docs[step['doc']].get_changes(step['have_deps']))
                self.results[step['res']] = out
            elif op == 'get_missing_deps':
                self.results[step['res']] = self.docs[step['doc']].get_missing_deps(step['heads'])
            elif op == 'blocks':
                if self.structural:
                    self.results[step['res']] = self.docs[step['doc']].blocks()
                else:
                    skip_refs.add(step['res'])
            elif op == 'bloom_contains':
                if step['bloom']['res'] in skip_refs:
                    skip_refs.add(step['res'])
                else:
                    self.results[step['res']] = bloom_contains(self.resolve(step['bloom']), step['actor'], step['ctr'])
            elif op == 'check_columns':
                if step['block']['res'] in skip_refs:
                    continue
                cols = self.resolve(step['block'])['columns']
                for name, hexv in step['expected'].items():
                    cid = DOC_OPS_COLUMNS.get(name, int(name) if name.isdigit() else -1)
                    if cid not in cols:
                        fails.append('checkColumns: missing column %s' % name)
                    elif cols[cid] != bytes.fromhex(hexv):
                        fails.append('checkColumns %s: %s != expected %s' % (name, list(cols[cid]), list(bytes.fromhex(hexv))))
                known = {v: k for k, v in DOC_OPS_COLUMNS.items()}
                for cid in cols:
                    name = known.get(cid, str(cid))
                    if name not in step['expected'] and name not in ('chldActor', 'chldCtr'):
                        fails.append('checkColumns: unexpected column %s' % name)
            elif op in ('assert_equal', 'assert_not_equal', 'assert_truthy'):
                sides = [step['actual']] + ([step['expected']] if 'expected' in step else [])
                if any('ref' in s and s['ref']['res'] in skip_refs for s in sides):
                    continue
                if any('ref' in s and s['ref']['res'] not in self.results for s in sides):
                    continue
                if op == 'assert_truthy':
                    if not self.resolve(step['actual']):
                        fails.append('assert(truthy) failed')
                    continue
                a, b = self.side(step['actual']), self.side(step['expected'])
                d = deep_equal(a, b)
                if op == 'assert_equal' and d:
                    fails.append('assert_equal: ' + d)
                if op == 'assert_not_equal' and not d:
                    fails.append('assert_not_equal: values are equal')
            # ---- Backend facade steps
            elif op == 'b_init':
                self.handles[step['h']] = self.facade.init()
            elif op == 'b_clone':
                out, failed = guarded(step, lambda: self.facade.clone(self.handles[step['src']]))
                if not failed:
                    self.handles[step['h']] = out
            elif op == 'b_free':
                self.facade.free(self.handles[step['h']])
            elif op == 'b_apply':
                out, failed = guarded(step, lambda: self.facade.applyChanges(self.handles[step['h']], changes_of(step)))
                if not failed:
                    self.handles[step['h2']] = out[0]
                    self.results[step['res']] = decode(out[1])
            elif op == 'b_apply_local':
                out, failed = guarded(step, lambda: self.facade.applyLocalChange(self.handles[step['h']], decode(step['change'])))
                if not failed:
                    self.handles[step['h2']] = out[0]
                    self.results[step['res']] = decode(out[1])
                    self.results[step['res_bin']] = out[2]
            elif op == 'b_save':
                out, failed = guarded(step, lambda: self.facade.save(self.handles[step['h']]))
                self.results[step['res']] = out
            elif op == 'b_load':
                if 'data' not in step or step['data'] is None:
                    fails.append('b_load: fixture carries no data')
                    continue
                out, failed = guarded(step, lambda: self.facade.load(bytes.fromhex(step['data'])))
                if not failed:
                    self.handles[step['h']] = out
            elif op == 'b_load_changes':
                out, failed = guarded(step, lambda: self.facade.loadChanges(self.handles[step['h']], changes_of(step)))
                if not failed:
                    self.handles[step['h2']] = out
            elif op == 'b_get_patch':
                out, failed = guarded(step, lambda: self.facade.getPatch(self.handles[step['h']]))
                self.results[step['res']] = decode(out)
            elif op == 'b_get_heads':
                self.results[step['res']] = self.facade.getHeads(self.handles[step['h']])
            elif op == 'b_get_all_changes':
                out, failed = guarded(step, lambda: self.facade.getAllChanges(self.handles[step['h']]))
                self.results[step['res']] = out
            elif op == 'b_get_changes':
                out, failed = guarded(step, lambda: self.facade.getChanges(self.handles[step['h']], decode(step['have_deps'])))
                self.results[step['res']] = out
            elif op == 'b_get_changes_added':
                out, failed = guarded(step, lambda: self.facade.getChangesAdded(self.handles[step['h1']], self.handles[step['h2']]))
                self.results[step['res']] = out
            elif op == 'b_get_change_by_hash':
                out, failed = guarded(step, lambda: self.facade.getChangeByHash(self.handles[step['h']], step['hash']))
                self.results[step['res']] = UNDEF if out is None else out
            elif op == 'b_get_missing_deps':
                out, failed = guarded(step, lambda: self.facade.getMissingDeps(self.handles[step['h']], step['heads']))
                self.results[step['res']] = out
            else:
                fails.append('unknown step ' + op)
        return fails
