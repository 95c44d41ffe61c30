// Flat binary patch (layout: include/amgpu.h, "Patch layout") -> the nested Patch object of
// @types/automerge/index.d.ts:236-316 that the frontend consumes (frontend/apply_patch.js). This is the JavaScript twin
// of automerge_classic_b200/engine.py FlatPatch.to_patch (which the parity tests of this repository drive): same passes,
// same rules. The records' keyOff / valOff index the patch buffer itself (its bytes section).
'use strict'

const ACTION_TYPE = {0: 'map', 2: 'list', 4: 'text', 6: 'table'}   // make* actions (columnar.js:52): even action numbers
const HDR_WORDS = 20, PROP_BYTES = 40, EDIT_BYTES = 32
const utf8 = new TextDecoder('utf-8')

function hex(bytes) { let s = ''; for (const b of bytes) s += (b < 16 ? '0' : '') + b.toString(16); return s }

// columnar.js:300-329 decodeValue: valLen = byteLength << 4 | type tag
function decodeValue(valLen, bytes) {
  const tag = valLen & 15
  if (valLen === 0) return {type: 'value', value: null}
  if (valLen === 1) return {type: 'value', value: false}
  if (valLen === 2) return {type: 'value', value: true}
  if (tag === 6) return {type: 'value', value: utf8.decode(bytes)}
  if (tag === 3 || tag === 4 || tag === 8 || tag === 9) {
    let v = 0n, shift = 0n, last = 0
    for (const b of bytes) { v |= BigInt(b & 0x7f) << shift; shift += 7n; last = b; if (!(b & 0x80)) break }
    if (tag !== 3 && (last & 0x40)) v -= 1n << shift
    return {type: 'value', value: Number(v), datatype: {3: 'uint', 4: 'int', 8: 'counter', 9: 'timestamp'}[tag]}
  }
  if (tag === 5) {
    if (bytes.byteLength !== 8) throw new RangeError(`Invalid length for floating point number: ${bytes.byteLength}`)
    return {type: 'value', value: new DataView(bytes.buffer, bytes.byteOffset, 8).getFloat64(0, true), datatype: 'float64'}
  }
  return {type: 'value', value: bytes.slice(), datatype: tag}
}

function emptyObjectPatch(objectId, action) {   // new.js:726-732
  const type = ACTION_TYPE[action]
  return (type === 'list' || type === 'text') ? {objectId, type, edits: []} : {objectId, type, props: {}}
}

function inflatePatch(buf, wholeDoc) {
  const view = new DataView(buf.buffer, buf.byteOffset, buf.byteLength)
  const u64 = i => Number(view.getBigUint64(8 * i, true))
  if (view.getBigUint64(0, true) !== 0x31504747414d41n) throw new Error('bad patch magic')
  const hdr = []; for (let i = 0; i < HDR_WORDS; i++) hdr.push(u64(i))
  const actors = []
  for (let i = 0, off = hdr[7]; i < hdr[8]; i++) {
    const len = view.getUint32(off, true)
    actors.push(hex(buf.subarray(off + 4, off + 4 + len)))
    off += 4 + len; off += (4 - off % 4) % 4
  }
  const opId = id => `${id >> 16n}@${actors[Number(id & 0xffffn)]}`
  const patch = {maxOp: hdr[1], clock: {}, deps: [], pendingChanges: hdr[2]}
  if (hdr[3]) { patch.actor = hex(buf.subarray(hdr[5], hdr[5] + hdr[6])); patch.seq = hdr[4] }
  for (let i = 0; i < hdr[10]; i++) patch.clock[actors[u64(hdr[9] / 8 + 2 * i)]] = u64(hdr[9] / 8 + 2 * i + 1)
  for (let i = 0; i < hdr[12]; i++) patch.deps.push(hex(buf.subarray(hdr[11] + 32 * i, hdr[11] + 32 * i + 32)))

  const prop = i => { const o = hdr[13] + PROP_BYTES * i; return {obj: view.getBigUint64(o, true), opId: view.getBigUint64(o + 8, true), keyOff: view.getUint32(o + 16, true), keyLen: view.getUint32(o + 20, true), valLen: view.getUint32(o + 24, true), valOff: view.getUint32(o + 28, true), flags: view.getUint32(o + 32, true), pad: view.getUint32(o + 36, true)} }
  const edit = i => { const o = hdr[15] + EDIT_BYTES * i; return {obj: view.getBigUint64(o, true), opId: view.getBigUint64(o + 8, true), index: view.getUint32(o + 16, true), kind: view.getUint32(o + 20, true), valLen: view.getUint32(o + 24, true), valOff: view.getUint32(o + 28, true)} }
  const elemId = i => hdr[17] ? view.getBigUint64(hdr[17] + 8 * i, true) : edit(i).opId   // section absent: every insert's elemId is its opId
  const int64 = (lo, hi) => Number(BigInt.asIntN(64, BigInt(lo) | (BigInt(hi) << 32n)))
  const nProps = hdr[14], nEdits = hdr[16]

  const patches = {_root: {objectId: '_root', type: 'map', props: {}}}
  // pass 1: every object that appears as a value gets its (empty) patch
  for (let i = 0; i < nProps; i++) { const r = prop(i), action = r.flags >>> 8; if (action % 2 === 0 && !(r.flags & 1)) { const id = opId(r.opId); if (!patches[id]) patches[id] = emptyObjectPatch(id, action) } }
  for (let i = 0; i < nEdits; i++) { const r = edit(i), action = r.kind >>> 16; if (action % 2 === 0 && (r.kind & 0xff) !== 1) { const id = opId(r.opId); if (!patches[id]) patches[id] = emptyObjectPatch(id, action) } }
  // pass 2: map entries
  for (let i = 0; i < nProps; i++) {
    const r = prop(i), p = patches[r.obj === 0n ? '_root' : opId(r.obj)]
    if (!p || !p.props) continue   // object not reachable from the root (its make op is no longer visible)
    const key = utf8.decode(buf.subarray(r.keyOff, r.keyOff + r.keyLen)), action = r.flags >>> 8
    const entry = () => (p.props[key] || (p.props[key] = {}))
    if (r.flags & 1) entry()                                                      // key without a visible value: `key: {}`
    else if (action === 1 && (r.flags & 2)) entry()[opId(r.opId)] = {type: 'value', value: int64(r.valOff, r.pad), datatype: 'counter'}   // increments summed by the engine (new.js:941-966)
    else if (action === 1) entry()[opId(r.opId)] = decodeValue(r.valLen, buf.subarray(r.valOff, r.valOff + (r.valLen >>> 4)))
    else if (action % 2 === 0) entry()[opId(r.opId)] = patches[opId(r.opId)]
    else if (!wholeDoc) entry()
  }
  // pass 3: list edits (ordered per object; 0x100 = starts a new run, 0x200 = rendered as multi-insert; new.js:747-782)
  for (let i = 0; i < nEdits; i++) {
    const r = edit(i), p = patches[opId(r.obj)]
    if (!p || !p.edits) continue
    const kind = r.kind & 0xff, runStart = !!(r.kind & 0x100), multi = !!(r.kind & 0x200), action = r.kind >>> 16, edits = p.edits
    if (kind === 1) { if (runStart) edits.push({action: 'remove', index: r.index, count: 1}); else edits[edits.length - 1].count += 1; continue }
    let value
    if (action === 1 && (r.kind & 0x1000)) value = {type: 'value', value: int64(r.valLen, r.valOff), datatype: 'counter'}
    else if (action === 1) value = decodeValue(r.valLen, buf.subarray(r.valOff, r.valOff + (r.valLen >>> 4)))
    else if (action % 2 === 0) value = patches[opId(r.opId)]
    else continue
    if (kind === 2) edits.push({action: 'update', index: r.index, opId: opId(r.opId), value})
    else if (!runStart) edits[edits.length - 1].values.push(value.value)        // continues the multi-insert opened by an earlier record
    else if (multi) { const e = {action: 'multi-insert', index: r.index, elemId: opId(elemId(i)), values: [value.value]}; if (value.datatype) e.datatype = value.datatype; edits.push(e) }
    else edits.push({action: 'insert', index: r.index, elemId: opId(elemId(i)), opId: opId(r.opId), value})
  }
  patch.diffs = patches._root
  return patch
}

module.exports = { inflatePatch, decodeValue }
