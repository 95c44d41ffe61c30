"""Host-side mirror of the reference's change *encoder* (backend/columnar.js encodeChange).

The reference keeps `encodeChange` in host JavaScript even with a native backend plugged in:
`Backend.applyLocalChange` (backend/backend.js:54-91) encodes the frontend's change request on the
host and then calls `applyChanges` with the binary change.  This module is that host-side encoder
for the Python mirror of the Backend facade (automerge_classic_b200/backend.py); it is also what the
fixture extractor (tests/jsfixtures/extract.py) uses to turn the reference tests' JSON changes into
bytes.  It is pinned byte-for-byte by the annotated golden change of test/columnar_test.js:8-37.

Follows (paths relative to /root/reference):
  backend/encoding.js:57-286     Encoder (LEB128)
  backend/encoding.js:558-783    RLEEncoder (as canonical batch encoder)
  backend/encoding.js:932-998    DeltaEncoder
  backend/encoding.js:1061-1135  BooleanEncoder
  backend/columnar.js:133-170    parseAllOpIds
  backend/columnar.js:176-292    encodeObjectId / encodeOperationKey / encodeOperationAction / encodeValue
  backend/columnar.js:370-436    encodeOps
  backend/columnar.js:446-475    expandMultiOps
  backend/columnar.js:659-686    encodeContainer
  backend/columnar.js:710-739    encodeChange
  src/common.js:22-28            parseOpId
"""
import hashlib
import re
import struct
import zlib

MAGIC = bytes([0x85, 0x6f, 0x4a, 0x83])
ACTIONS = ['makeMap', 'set', 'makeList', 'del', 'makeText', 'inc', 'makeTable', 'link']
VT = dict(NULL=0, FALSE=1, TRUE=2, LEB128_UINT=3, LEB128_INT=4, IEEE754=5, UTF8=6, BYTES=7, COUNTER=8, TIMESTAMP=9)
MAX_SAFE = 2 ** 53 - 1
DEFLATE_MIN_SIZE = 256


def uleb(v):
    if v < 0 or v > MAX_SAFE:
        raise ValueError('number out of range')
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def sleb(v):
    if v < -MAX_SAFE or v > MAX_SAFE:
        raise ValueError('number out of range')
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        if (v == 0 and not b & 0x40) or (v == -1 and b & 0x40):
            out.append(b)
            return bytes(out)
        out.append(b | 0x80)


def prefixed(b):
    return uleb(len(b)) + b


def hex_bytes(h):
    if not re.fullmatch(r'([0-9a-f][0-9a-f])*', h):
        raise ValueError('value is not hexadecimal')
    return bytes.fromhex(h)


def rle_encode(values, kind):
    """kind: 'uint' | 'int' | 'utf8'; values may contain None."""
    def raw(v):
        if kind == 'uint':
            return uleb(v)
        if kind == 'int':
            return sleb(v)
        return prefixed(v.encode('utf-8'))
    if all(v is None for v in values):
        return b''
    runs = []
    for v in values:
        if runs and runs[-1][0] == v and type(runs[-1][0]) is type(v):
            runs[-1][1] += 1
        else:
            runs.append([v, 1])
    out = bytearray()
    lit = []

    def flush():
        if lit:
            out.extend(sleb(-len(lit)))
            for x in lit:
                out.extend(raw(x))
            lit.clear()
    for v, n in runs:
        if v is None:
            flush()
            out.extend(sleb(0))
            out.extend(uleb(n))
        elif n >= 2:
            flush()
            out.extend(sleb(n))
            out.extend(raw(v))
        else:
            lit.append(v)
    flush()
    return bytes(out)


def delta_encode(values):
    out, last = [], 0
    for v in values:
        if v is None:
            out.append(None)
        else:
            out.append(v - last)
            last = v
    return rle_encode(out, 'int')


def bool_encode(values):
    out = bytearray()
    last, count = False, 0
    for v in values:
        if v == last:
            count += 1
        else:
            out.extend(uleb(count))
            last, count = v, 1
    if count > 0:
        out.extend(uleb(count))
    return bytes(out)


def parse_op_id(s):
    m = re.fullmatch(r'(\d+)@(.*)', s or '')
    if not m:
        raise ValueError('Not a valid opId: %s' % s)
    return int(m.group(1)), m.group(2)


def expand_multi_ops(ops, start_op, actor):
    op_num, out = start_op, []
    for op in ops:
        if op.get('action') == 'set' and 'values' in op and op.get('insert'):
            if op.get('pred'):
                raise ValueError('multi-insert pred must be empty')
            last = op['elemId']
            for value in op['values']:
                dt = op.get('datatype')
                ok = (isinstance(value, (str, bool)) or value is None) if dt is None else (isinstance(value, (int, float)) and not isinstance(value, bool))
                if not ok:
                    raise ValueError('Decode failed: bad value/datatype association (%s,%s)' % (value, dt))
                o = {'action': 'set', 'obj': op['obj'], 'elemId': last, 'value': value, 'pred': [], 'insert': True}
                if 'datatype' in op:
                    o['datatype'] = op['datatype']
                out.append(o)
                last = '%d@%s' % (op_num, actor)
                op_num += 1
        elif op.get('action') == 'del' and op.get('multiOp', 0) > 1:
            if len(op['pred']) != 1:
                raise ValueError('multiOp deletion must have exactly one pred')
            ec, ea = parse_op_id(op['elemId'])
            pc, pa = parse_op_id(op['pred'][0])
            for i in range(op['multiOp']):
                out.append({'action': 'del', 'obj': op['obj'], 'elemId': '%d@%s' % (ec + i, ea), 'pred': ['%d@%s' % (pc + i, pa)]})
                op_num += 1
        else:
            out.append(op)
            op_num += 1
    return out


def encode_value(op, val_len, val_raw):
    action, value = op.get('action'), op.get('value')
    if action not in ('set', 'inc') or value is None:
        val_len.append(VT['NULL'])
    elif value is False:
        val_len.append(VT['FALSE'])
    elif value is True:
        val_len.append(VT['TRUE'])
    elif isinstance(value, str):
        b = value.encode('utf-8')
        val_raw.extend(b)
        val_len.append(len(b) << 4 | VT['UTF8'])
    elif isinstance(value, (bytes, bytearray)):
        dt = op.get('datatype')
        tag = dt if isinstance(dt, int) and 10 <= dt <= 15 else VT['BYTES']
        val_raw.extend(value)
        val_len.append(len(value) << 4 | tag)
    elif isinstance(value, (int, float)):
        dt = op.get('datatype')
        if dt == 'counter':
            tag, b = VT['COUNTER'], sleb(int(value))
        elif dt == 'timestamp':
            tag, b = VT['TIMESTAMP'], sleb(int(value))
        elif dt == 'uint':
            tag, b = VT['LEB128_UINT'], uleb(int(value))
        elif dt == 'int':
            tag, b = VT['LEB128_INT'], sleb(int(value))
        elif dt == 'float64':
            tag, b = VT['IEEE754'], struct.pack('<d', float(value))
        elif float(value).is_integer() and abs(value) <= MAX_SAFE and not (isinstance(value, float) and dt == 'float64'):
            tag, b = VT['LEB128_INT'], sleb(int(value))
        else:
            tag, b = VT['IEEE754'], struct.pack('<d', float(value))
        val_raw.extend(b)
        val_len.append(len(b) << 4 | tag)
    else:
        raise ValueError('Unsupported value in operation: %r' % (value,))


def encode_change_raw(change, compress=True, level=6):
    """Returns (bytes, hash_hex). `change` is the JSON form used throughout the reference's tests.
    `level` is the zlib level of the DEFLATE step (the reference uses pako's default, 6)."""
    actor = change['actor']
    ops = expand_multi_ops(change['ops'], change['startOp'], actor)
    # parseAllOpIds(single=True): actor table = [author] + sorted(other actors)
    actors = {actor}
    parsed = []
    for op in ops:
        p = dict(op)
        p['_obj'] = None if op['obj'] == '_root' else parse_op_id(op['obj'])
        elem = op.get('elemId')
        p['_elem'] = elem if (elem is None or elem == '_head') else parse_op_id(elem)
        p['_child'] = parse_op_id(op['child']) if op.get('child') else None
        p['_pred'] = [parse_op_id(x) for x in op.get('pred', [])]
        for ref in [p['_obj'], p['_elem'] if isinstance(p['_elem'], tuple) else None, p['_child']] + p['_pred']:
            if ref:
                actors.add(ref[1])
        parsed.append(p)
    actor_ids = [actor] + sorted(a for a in actors if a != actor)
    num = {a: i for i, a in enumerate(actor_ids)}

    cols = {k: [] for k in ['objActor', 'objCtr', 'keyActor', 'keyCtr', 'keyStr', 'insert', 'action', 'valLen',
                            'chldActor', 'chldCtr', 'predNum', 'predActor', 'predCtr']}
    val_raw = bytearray()
    for p in parsed:
        if p['_obj'] is None:
            cols['objActor'].append(None); cols['objCtr'].append(None)
        else:
            if p['_obj'][0] <= 0:
                raise ValueError('Unexpected objectId reference')
            cols['objActor'].append(num[p['_obj'][1]]); cols['objCtr'].append(p['_obj'][0])
        if p.get('key'):
            cols['keyActor'].append(None); cols['keyCtr'].append(None); cols['keyStr'].append(p['key'])
        elif p['_elem'] == '_head' and p.get('insert'):
            cols['keyActor'].append(None); cols['keyCtr'].append(0); cols['keyStr'].append(None)
        elif isinstance(p['_elem'], tuple) and p['_elem'][0] > 0:
            cols['keyActor'].append(num[p['_elem'][1]]); cols['keyCtr'].append(p['_elem'][0]); cols['keyStr'].append(None)
        else:
            raise ValueError('Unexpected operation key: %r' % (p,))
        cols['insert'].append(bool(p.get('insert')))
        a = p['action']
        if a in ACTIONS:
            cols['action'].append(ACTIONS.index(a))
        elif isinstance(a, int):
            cols['action'].append(a)
        else:
            raise ValueError('Unexpected operation action: %r' % (a,))
        encode_value(p, cols['valLen'], val_raw)
        if p['_child'] and p['_child'][0]:
            cols['chldActor'].append(num[p['_child'][1]]); cols['chldCtr'].append(p['_child'][0])
        else:
            cols['chldActor'].append(None); cols['chldCtr'].append(None)
        preds = sorted(p['_pred'])   # compareParsedOpIds: counter, then actorId string
        cols['predNum'].append(len(preds))
        for c, a_ in preds:
            cols['predActor'].append(num[a_]); cols['predCtr'].append(c)

    column_list = [
        (0x01, rle_encode(cols['objActor'], 'uint')), (0x02, rle_encode(cols['objCtr'], 'uint')),
        (0x11, rle_encode(cols['keyActor'], 'uint')), (0x13, delta_encode(cols['keyCtr'])),
        (0x15, rle_encode(cols['keyStr'], 'utf8')), (0x34, bool_encode(cols['insert'])),
        (0x42, rle_encode(cols['action'], 'uint')), (0x56, rle_encode(cols['valLen'], 'uint')),
        (0x57, bytes(val_raw)), (0x61, rle_encode(cols['chldActor'], 'uint')), (0x63, delta_encode(cols['chldCtr'])),
        (0x70, rle_encode(cols['predNum'], 'uint')), (0x71, rle_encode(cols['predActor'], 'uint')),
        (0x73, delta_encode(cols['predCtr']))]

    body = bytearray()
    deps = sorted(change.get('deps', []))
    body += uleb(len(deps))
    for d in deps:
        body += hex_bytes(d)
    body += prefixed(hex_bytes(actor)) + uleb(change['seq']) + uleb(change['startOp']) + sleb(change.get('time', 0))
    body += prefixed((change.get('message') or '').encode('utf-8'))
    body += uleb(len(actor_ids) - 1)
    for a in actor_ids[1:]:
        body += prefixed(hex_bytes(a))
    nonempty = [(cid, buf) for cid, buf in column_list if len(buf) > 0]
    body += uleb(len(nonempty))
    for cid, buf in nonempty:
        body += uleb(cid) + uleb(len(buf))
    for cid, buf in column_list:
        body += buf
    if change.get('extraBytes'):
        body += bytes(change['extraBytes'])

    header = bytes([1]) + uleb(len(body))
    digest = hashlib.sha256(header + bytes(body)).digest()
    raw = MAGIC + digest[:4] + header + bytes(body)
    if compress and len(raw) >= DEFLATE_MIN_SIZE:
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = co.compress(bytes(body)) + co.flush()
        raw = MAGIC + digest[:4] + bytes([2]) + uleb(len(comp)) + comp
    return raw, digest.hex()


def encode_change(change, compress=True, level=6):
    return encode_change_raw(change, compress, level)[0]


def change_hash(change):
    return encode_change_raw(change, False)[1]


# ---------------------------------------------------------------------------------------------
# Decode side (host mirror of decodeChange; used by tests and by callers that want JSON changes).
# Follows backend/encoding.js:293-534 (Decoder), :789-920 (RLEDecoder), :1004-1051 (DeltaDecoder),
# :1141-1207 (BooleanDecoder); backend/columnar.js:300-361 (decodeValue, decodeValueColumns),
# :483-523 (decodeOps, checkSortedOpIds), :577-607 (decodeColumns), :609-652, :688-708, :741-776.

class DecodeError(ValueError):
    """The reference throws RangeError for all of these."""


class _Dec:
    def __init__(self, buf):
        self.buf, self.off = bytes(buf), 0

    @property
    def done(self):
        return self.off == len(self.buf)

    def uleb(self, limit=MAX_SAFE):
        result, shift = 0, 0
        while self.off < len(self.buf):
            b = self.buf[self.off]
            if shift == 63 and b & 0xfe:
                raise DecodeError('number out of range')
            result |= (b & 0x7f) << shift
            shift += 7
            self.off += 1
            if not b & 0x80:
                if result > limit:
                    raise DecodeError('number out of range')
                return result
        raise DecodeError('buffer ended with incomplete number')

    def sleb(self):
        result, shift = 0, 0
        while self.off < len(self.buf):
            b = self.buf[self.off]
            if shift == 63 and b not in (0, 0x7f):
                raise DecodeError('number out of range')
            result |= (b & 0x7f) << shift
            shift += 7
            self.off += 1
            if not b & 0x80:
                if b & 0x40:
                    result -= 1 << shift
                if result < -MAX_SAFE or result > MAX_SAFE:
                    raise DecodeError('number out of range')
                return result
        raise DecodeError('buffer ended with incomplete number')

    def raw(self, n):
        if self.off + n > len(self.buf):
            raise DecodeError('subarray exceeds buffer size')
        self.off += n
        return self.buf[self.off - n:self.off]

    def prefixed(self):
        return self.raw(self.uleb())


def rle_decode(buf, kind):
    d, out, state, last = _Dec(buf), [], None, object()
    def raw():
        if kind == 'uint':
            return d.uleb()
        if kind == 'int':
            return d.sleb()
        return d.prefixed().decode('utf-8', 'replace')
    while not d.done:
        count = d.sleb()
        if count > 1:
            v = raw()
            if state in ('rep', 'lit') and v == last:
                raise DecodeError('Successive repetitions with the same value are not allowed')
            state, last = 'rep', v
            out.extend([v] * count)
        elif count == 1:
            raise DecodeError('Repetition count of 1 is not allowed, use a literal instead')
        elif count < 0:
            if state == 'lit':
                raise DecodeError('Successive literals are not allowed')
            state = 'lit'
            for _ in range(-count):
                v = raw()
                if v == last:
                    raise DecodeError('Repetition of values is not allowed in literal')
                last = v
                out.append(v)
        else:
            if state == 'nulls':
                raise DecodeError('Successive null runs are not allowed')
            n = d.uleb()
            if n == 0:
                raise DecodeError('Zero-length null runs are not allowed')
            state, last = 'nulls', None
            out.extend([None] * n)
    return out


def delta_decode(buf):
    out, acc = [], 0
    for v in rle_decode(buf, 'int'):
        if v is None:
            out.append(None)
        else:
            acc += v
            out.append(acc)
    return out


def bool_decode(buf):
    d, out, val, first = _Dec(buf), [], True, True
    while not d.done:
        n = d.uleb()
        val = not val
        if n == 0 and not first:
            raise DecodeError('Zero-length runs are not allowed')
        first = False
        out.extend([val] * n)
    return out


def decode_value(size_tag, raw):
    tag = size_tag % 16
    if size_tag == VT['NULL']:
        return None, None
    if size_tag == VT['FALSE']:
        return False, None
    if size_tag == VT['TRUE']:
        return True, None
    if tag == VT['UTF8']:
        return raw.decode('utf-8', 'replace'), None
    if tag == VT['LEB128_UINT']:
        return _Dec(raw).uleb(), 'uint'
    if tag == VT['LEB128_INT']:
        return _Dec(raw).sleb(), 'int'
    if tag == VT['IEEE754']:
        if len(raw) != 8:
            raise DecodeError('Invalid length for floating point number: %d' % len(raw))
        return struct.unpack('<d', raw)[0], 'float64'
    if tag == VT['COUNTER']:
        return _Dec(raw).sleb(), 'counter'
    if tag == VT['TIMESTAMP']:
        return _Dec(raw).sleb(), 'timestamp'
    return bytes(raw), tag


def split_container(buf):
    """columnar.js:688-708 -> (chunk_type, body, hash_hex); verifies magic + checksum."""
    buf = bytes(buf)
    d = _Dec(buf)
    if d.raw(4) != MAGIC:
        raise DecodeError('Data does not begin with magic bytes 85 6f 4a 83')
    expected = d.raw(4)
    start = d.off
    chunk_type = d.raw(1)[0]
    body = d.raw(d.uleb())
    digest = hashlib.sha256(buf[start:d.off]).digest()
    if digest[:4] != expected:
        raise DecodeError('checksum does not match data')
    return chunk_type, body, digest.hex(), d.off


def inflate_change(buf):
    buf = bytes(buf)
    if len(buf) > 8 and buf[8] == 2:
        d = _Dec(buf[9:])
        comp = d.raw(d.uleb())
        body = zlib.decompress(comp, -15)
        return buf[:8] + bytes([1]) + uleb(len(body)) + body
    return buf


def _check_sorted(ids):
    last = None
    for cur in ids:
        if last is not None and not (last[0] < cur[0] or (last[0] == cur[0] and last[1] < cur[1])):
            raise DecodeError('operation IDs are not in ascending order')
        last = cur


def decode_change(buf):
    """columnar.js:770-776 decodeChange: binary change -> JSON change (with `hash`)."""
    buf = inflate_change(buf)
    chunk_type, body, hash_hex, end = split_container(buf)
    if end != len(buf):
        raise DecodeError('Encoded change has trailing data')
    if chunk_type != 1:
        raise DecodeError('Unexpected chunk type: %d' % chunk_type)
    d = _Dec(body)
    deps = [d.raw(32).hex() for _ in range(d.uleb())]
    actor = d.prefixed().hex()
    change = {'actor': actor, 'seq': d.uleb(), 'startOp': d.uleb(), 'time': d.sleb(),
              'message': d.prefixed().decode('utf-8', 'replace'), 'deps': deps}
    actor_ids = [actor] + [d.prefixed().hex() for _ in range(d.uleb())]
    infos, last_id = [], -1
    for _ in range(d.uleb()):
        cid, ln = d.uleb(), d.uleb()
        if last_id >= 0 and (cid & ~8) <= (last_id & ~8):
            raise DecodeError('Columns must be in ascending order')
        last_id = cid
        infos.append((cid, ln))
    cols = {}
    for cid, ln in infos:
        if cid & 8:
            raise DecodeError('change must not contain deflated columns')
        cols[cid] = d.raw(ln)
    if not d.done:
        change['extraBytes'] = d.raw(len(d.buf) - d.off)

    def col(cid, kind):
        b = cols.get(cid, b'')
        return bool_decode(b) if kind == 'bool' else (delta_decode(b) if kind == 'delta' else rle_decode(b, kind))
    action = col(0x42, 'uint')
    n = len(action)
    for cid, b in cols.items():   # the row count is the longest column (decodeColumns: any column not done)
        if cid not in (0x57, 0x71, 0x73) and cid in (0x01, 0x02, 0x11, 0x13, 0x15, 0x34, 0x56, 0x61, 0x63, 0x70):
            kind = 'bool' if cid == 0x34 else ('delta' if cid & 7 == 3 else ('utf8' if cid & 7 == 5 else 'uint'))
            n = max(n, len(col(cid, kind)))

    def padded(values, fill=None):
        return values + [fill] * (n - len(values))
    obj_actor, obj_ctr = padded(col(0x01, 'uint')), padded(col(0x02, 'uint'))
    key_actor, key_ctr, key_str = padded(col(0x11, 'uint')), padded(col(0x13, 'delta')), padded(col(0x15, 'utf8'))
    insert, action, val_len = padded(col(0x34, 'bool'), False), padded(action), padded(col(0x56, 'uint'))
    chld_actor, chld_ctr = padded(col(0x61, 'uint')), padded(col(0x63, 'delta'))
    pred_num, pred_actor, pred_ctr = padded(col(0x70, 'uint')), col(0x71, 'uint'), col(0x73, 'delta')
    raw = _Dec(cols.get(0x57, b''))

    def actor_of(i):
        if i is None:
            return None
        if i >= len(actor_ids):
            raise DecodeError('No actor index %d' % i)
        return actor_ids[i]
    ops, pi = [], 0
    for i in range(n):
        obj = '_root' if obj_ctr[i] is None else '%d@%s' % (obj_ctr[i], actor_of(obj_actor[i]))
        act = ACTIONS[action[i]] if action[i] is not None and action[i] < len(ACTIONS) else action[i]
        if key_str[i]:
            op = {'obj': obj, 'key': key_str[i], 'action': act}
        else:
            op = {'obj': obj, 'elemId': '_head' if key_ctr[i] == 0 else '%s@%s' % (key_ctr[i], actor_of(key_actor[i])), 'action': act}
        op['insert'] = bool(insert[i])
        size_tag = val_len[i] if val_len[i] is not None else 0
        value, datatype = decode_value(size_tag, raw.raw(size_tag >> 4))
        if act in ('set', 'inc'):
            op['value'] = value
            if datatype:
                op['datatype'] = datatype
        if bool(chld_ctr[i]) != bool(chld_actor[i] is not None and actor_of(chld_actor[i])):
            raise DecodeError('Mismatched child columns: %s and %s' % (chld_ctr[i], chld_actor[i]))
        if chld_ctr[i] is not None:
            op['child'] = '%d@%s' % (chld_ctr[i], actor_of(chld_actor[i]))
        k = pred_num[i] or 0
        preds = [(pred_ctr[pi + j] if pi + j < len(pred_ctr) else None,
                  actor_of(pred_actor[pi + j]) if pi + j < len(pred_actor) else None) for j in range(k)]
        pi += k
        _check_sorted(preds)
        op['pred'] = ['%s@%s' % p for p in preds]
        ops.append(op)
    change['ops'] = ops
    change['hash'] = hash_hex
    return change
