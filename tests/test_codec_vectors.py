"""Known-answer vectors of the reference's codec tests (test/encoding_test.js:521-604 RLE, :764-785 delta,
:908-957 boolean, LEB128 samples of :7-120 / :122-256; test/columnar_test.js golden changes), against
(a) the C++ oracle codecs and (b) the host-side Python mirror automerge_classic_b200/columnar.py."""
import pytest

from automerge_classic_b200 import columnar as col

RLE = [  # (kind, values, bytes)  test/encoding_test.js:521-566
    ('uint', [], []), ('uint', [1, 2, 3], [0x7d, 1, 2, 3]), ('uint', [0, 1, 2, 2, 3], [0x7e, 0, 1, 2, 2, 0x7f, 3]),
    ('uint', [1, 1, 1, 1, 1, 1], [6, 1]), ('uint', [1, 1, 1, 4, 4, 4], [3, 1, 3, 4]), ('uint', [0xff], [0x7f, 0xff, 0x01]),
    ('int', [-0x40], [0x7f, 0x40]),
    ('uint', [None, 1], [0, 1, 0x7f, 1]), ('uint', [1, None], [0x7f, 1, 0, 1]), ('uint', [1, 1, 1, None], [3, 1, 0, 1]),
    ('uint', [None, None, None, 3, 4, 5, None], [0, 3, 0x7d, 3, 4, 5, 0, 1]), ('uint', [None, None, None, 9, 9, 9], [0, 3, 3, 9]),
    ('uint', [1, 1, 1, 1, 1, None, None, None, 1], [5, 1, 0, 3, 0x7f, 1]),
    ('utf8', ['a'], [0x7f, 1, 0x61]), ('utf8', ['a', 'b', 'c', 'd'], [0x7c, 1, 0x61, 1, 0x62, 1, 0x63, 1, 0x64]),
    ('utf8', ['a', 'a', 'a', 'a'], [4, 1, 0x61]), ('utf8', ['a', 'a', None, None, 'a', 'a'], [2, 1, 0x61, 0, 2, 2, 1, 0x61]),
    ('utf8', [None, None, None, None, 'abc'], [0, 4, 0x7f, 3, 0x61, 0x62, 0x63]),
    ('uint', [None], []), ('uint', [None, None, None, None], []),
]
DELTA = [  # test/encoding_test.js:764-772
    ([], []), ([18, 2, 9, 15, 16, 19, 25], [0x79, 18, 0x70, 7, 6, 1, 3, 6]), ([1, 2, 3, 4, 5, 6, 7, 8], [8, 1]),
    ([10, 11, 12, 13, 14, 15], [0x7f, 10, 5, 1]), ([10, 11, 12, 13, 0, 1, 2, 3], [0x7f, 10, 3, 1, 0x7f, 0x73, 3, 1]),
    ([0, 1, 2, 3, None, None, None, 4, 5, 6], [0x7f, 0, 3, 1, 0, 3, 3, 1]), ([-64, -60, -56, -52, -48, -44, -40, -36], [0x7f, 0x40, 7, 4]),
]
BOOL = [  # test/encoding_test.js:908-915
    ([], []), ([False], [1]), ([True], [0, 1]), ([False, False, False, True, True], [3, 2]), ([True, True, True, False, False], [0, 3, 2]),
    ([True, False, True, False, True, True, False], [0, 1, 1, 1, 1, 2, 1]),
]
RLE_ERRORS = [  # test/encoding_test.js:594-603
    ([1, 1], 'Repetition count of 1 is not allowed'), ([2, 1, 2, 1], 'Successive repetitions with the same value'),
    ([0, 1, 0, 2], 'Successive null runs are not allowed'), ([0, 0], 'Zero-length null runs are not allowed'),
    ([0x7f, 1, 0x7f, 2], 'Successive literals are not allowed'), ([0x7d, 1, 2, 2], 'Repetition of values is not allowed'),
    ([2, 0, 0x7e, 0, 1], 'Repetition of values is not allowed'), ([0x7e, 1, 2, 2, 2], 'Successive repetitions with the same value'),
]
LEB_U = [(0, [0]), (1, [1]), (0x42, [0x42]), (0x7f, [0x7f]), (0x80, [0x80, 0x01]), (0xff, [0xff, 0x01]), (0x1234, [0xb4, 0x24]),
         (0x3fff, [0xff, 0x7f]), (0x4000, [0x80, 0x80, 0x01]), (0xffffffff, [0xff, 0xff, 0xff, 0xff, 0x0f]),
         (2 ** 53 - 1, [0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0x0f])]
LEB_S = [(0, [0]), (1, [1]), (-1, [0x7f]), (0x3f, [0x3f]), (0x40, [0xc0, 0x00]), (-0x3f, [0x41]), (-0x40, [0x40]), (-0x41, [0xbf, 0x7f]),
         (0x1fff, [0xff, 0x3f]), (0x2000, [0x80, 0xc0, 0x00]), (-0x2000, [0x80, 0x40]), (-0x2001, [0xff, 0xbf, 0x7f]),
         (2 ** 53 - 1, [0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0x0f]), (-(2 ** 53 - 1), [0x81, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80, 0x70])]


@pytest.mark.parametrize('kind,values,expected', RLE)
def test_rle(oracle_mod, kind, values, expected):
    assert list(oracle_mod.encode_column(kind, values)) == expected
    assert oracle_mod.decode_column(kind, bytes(expected)) == ([] if not expected else values)
    assert list(col.rle_encode(values, kind)) == expected
    assert col.rle_decode(bytes(expected), kind) == ([] if not expected else values)


@pytest.mark.parametrize('values,expected', DELTA)
def test_delta(oracle_mod, values, expected):
    assert list(oracle_mod.encode_column('delta', values)) == expected
    assert oracle_mod.decode_column('delta', bytes(expected)) == values
    assert list(col.delta_encode(values)) == expected and col.delta_decode(bytes(expected)) == values


@pytest.mark.parametrize('values,expected', BOOL)
def test_boolean(oracle_mod, values, expected):
    assert list(oracle_mod.encode_column('boolean', [int(v) for v in values])) == expected
    assert oracle_mod.decode_column('boolean', bytes(expected)) == values
    assert list(col.bool_encode(values)) == expected and col.bool_decode(bytes(expected)) == values


@pytest.mark.parametrize('data,message', RLE_ERRORS)
def test_rle_canonical_form(oracle_mod, data, message):
    with pytest.raises(oracle_mod.OracleError, match=message):
        oracle_mod.decode_column('int', bytes(data))
    with pytest.raises(col.DecodeError, match=message):
        col.rle_decode(bytes(data), 'int')


def test_boolean_canonical_form(oracle_mod):
    for data in ([1, 0], [1, 1, 0]):
        with pytest.raises(oracle_mod.OracleError, match='Zero-length runs are not allowed'):
            oracle_mod.decode_column('boolean', bytes(data))
        with pytest.raises(col.DecodeError, match='Zero-length runs are not allowed'):
            col.bool_decode(bytes(data))


def test_leb128(oracle_mod):
    for v, b in LEB_U:
        assert list(oracle_mod.leb_encode('uint53', v)) == b and oracle_mod.leb_decode('uint53', bytes(b)) == (v, len(b))
        assert list(col.uleb(v)) == b
    for v, b in LEB_S:
        assert list(oracle_mod.leb_encode('int53', v)) == b and oracle_mod.leb_decode('int53', bytes(b)) == (v, len(b))
        assert list(col.sleb(v)) == b
    with pytest.raises(oracle_mod.OracleError, match='number out of range'):
        oracle_mod.leb_decode('uint53', bytes([0xff] * 7 + [0x1f]))        # 2^53 and above
    with pytest.raises(oracle_mod.OracleError, match='incomplete number'):
        oracle_mod.leb_decode('uint53', bytes([0x80, 0x80]))


GOLDEN_CHANGE = bytes([  # test/columnar_test.js:15-37, every byte annotated there
    0x85, 0x6f, 0x4a, 0x83, 0xe2, 0xbd, 0xfb, 0xf5, 1, 94, 0, 2, 0xaa, 0xaa, 1, 1, 9, 0, 0, 12, 0x01, 4, 0x02, 4, 0x11, 8, 0x13, 7, 0x15, 8,
    0x34, 4, 0x42, 6, 0x56, 6, 0x57, 3, 0x70, 6, 0x71, 2, 0x73, 2, 0, 1, 4, 0, 0, 1, 4, 1, 0, 2, 0x7f, 0, 0, 1, 0x7f, 0, 0, 1, 0x7c, 0, 2, 0x7e, 4,
    0x7f, 4, 0x74, 0x65, 0x78, 0x74, 0, 4, 1, 1, 1, 2, 0x7d, 4, 1, 3, 2, 1, 0x7d, 0, 0x16, 0, 2, 0x16, 0x68, 0x48, 0x69, 2, 0, 0x7f, 1, 2, 0, 0x7f, 0, 0x7f, 2])
GOLDEN_JSON = {'actor': 'aaaa', 'seq': 1, 'startOp': 1, 'time': 9, 'message': '', 'deps': [], 'ops': [
    {'action': 'makeText', 'obj': '_root', 'key': 'text', 'insert': False, 'pred': []},
    {'action': 'set', 'obj': '1@aaaa', 'elemId': '_head', 'insert': True, 'value': 'h', 'pred': []},
    {'action': 'del', 'obj': '1@aaaa', 'elemId': '2@aaaa', 'insert': False, 'pred': ['2@aaaa']},
    {'action': 'set', 'obj': '1@aaaa', 'elemId': '_head', 'insert': True, 'value': 'H', 'pred': []},
    {'action': 'set', 'obj': '1@aaaa', 'elemId': '4@aaaa', 'insert': True, 'value': 'i', 'pred': []}]}
UNSORTED_PREDS = bytes([  # test/columnar_test.js:42-52
    133, 111, 74, 131, 31, 229, 112, 44, 1, 105, 1, 58, 30, 190, 100, 253, 180, 180, 66, 49, 126, 81, 142, 10, 3, 35, 140, 189, 231, 34, 145, 57, 66, 23, 224,
    149, 64, 97, 88, 140, 168, 194, 229, 4, 244, 209, 58, 138, 67, 140, 1, 152, 236, 250, 2, 0, 1, 4, 55, 234, 66, 242, 8, 21, 11, 52, 1, 66, 2, 86, 3, 87, 10, 112,
    2, 113, 3, 115, 4, 127, 9, 99, 111, 109, 109, 111, 110, 86, 97, 114, 1, 127, 1, 127, 166, 1, 52, 48, 57, 49, 52, 57, 52, 53, 56, 50, 127, 2, 126, 0, 1, 126, 139, 1, 0])
TRAILING = bytes([  # test/columnar_test.js:55-71
    0x85, 0x6f, 0x4a, 0x83, 0xb2, 0x98, 0x9e, 0xa9, 1, 61, 0, 2, 0x12, 0x34, 1, 1, 252, 250, 220, 255, 5, 14, 73, 110, 105, 116, 105, 97, 108, 105, 122, 97, 116, 105, 111, 110,
    0, 6, 0x15, 3, 0x34, 1, 0x42, 2, 0x56, 2, 0x57, 1, 0x70, 2, 0x7f, 1, 0x78, 1, 0x7f, 1, 0x7f, 19, 1, 0x7f, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9])


def test_golden_change(oracle_mod):
    assert col.encode_change(GOLDEN_JSON) == GOLDEN_CHANGE
    dec = col.decode_change(GOLDEN_CHANGE)
    assert dec['hash'].startswith('e2bdfbf5') and {k: dec[k] for k in GOLDEN_JSON} == GOLDEN_JSON
    o = oracle_mod.decode_change(GOLDEN_CHANGE)
    assert o['hash'] == dec['hash'] and len(o['ops']) == 5 and o['ops'][2]['pred'] == [[2, 0]]
    assert oracle_mod.sha256(GOLDEN_CHANGE[8:])[:4] == GOLDEN_CHANGE[4:8]


def test_unsorted_preds_rejected_by_decode_change_only(oracle_mod):
    with pytest.raises(col.DecodeError, match='operation IDs are not in ascending order'):
        col.decode_change(UNSORTED_PREDS)
    assert len(oracle_mod.decode_change(UNSORTED_PREDS)['ops']) == 1   # the hot path does not check pred order


def test_trailing_bytes_round_trip(oracle_mod):
    dec = col.decode_change(TRAILING)
    assert dec['extraBytes'] == bytes(range(10)) and dec['message'] == 'Initialization'
    assert col.encode_change(dec) == TRAILING
    assert oracle_mod.decode_change(TRAILING)['extraBytes'] == bytes(range(10)).hex()


def test_golden_hash_backend_test():
    # test/backend_test.js:735
    c = {'actor': '111111', 'seq': 1, 'time': 0, 'startOp': 1, 'deps': [], 'ops': [{'action': 'set', 'obj': '_root', 'key': 'bird', 'value': 'magpie', 'pred': []}]}
    assert col.change_hash(c) == '2c2845859ce4336936f56410f9161a09ba269f48aee5826782f1c389ec01d054'
