// ORACLE — TEST INFRASTRUCTURE ONLY (see codec.hpp header).
// extern "C" surface of the CPU oracle for ctypes (tests/, smoke(), bench.py cpu_baseline only).
#include <cstdio>
#include <cstdlib>
#include <sstream>
#include "backend.hpp"

using namespace orc;

namespace {

void jsonStr(std::string& out, const std::string& s) {
  out.push_back('"');
  for (unsigned char c : s) {
    switch (c) {
      case '"': out += "\\\""; break; case '\\': out += "\\\\"; break; case '\n': out += "\\n"; break;
      case '\r': out += "\\r"; break; case '\t': out += "\\t"; break;
      default: if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); out += b; } else out.push_back((char)c);
    }
  }
  out.push_back('"');
}
void jsonPrimValue(std::string& out, const Prim& p) {
  switch (p.k) {
    case Prim::P_NULL: out += "null"; break;
    case Prim::P_BOOL: out += p.b ? "true" : "false"; break;
    case Prim::P_INT: out += std::to_string(p.i); break;
    case Prim::P_FLOAT: {
      if (p.f != p.f) out += "NaN"; else if (p.f > 1.7976931348623157e308) out += "Infinity"; else if (p.f < -1.7976931348623157e308) out += "-Infinity";
      else { char b[40]; snprintf(b, sizeof b, "%.17g", p.f); out += b; if (!strpbrk(b, ".eEn")) out += ".0"; }
      break; }
    case Prim::P_STR: jsonStr(out, p.s); break;
    case Prim::P_BYTES: out += "{\"$bytes\":\""; out += toHex(p.s); out += "\"}"; break;
  }
}
void jsonDatatype(std::string& out, const std::string& dt, int dtNum) {
  if (!dt.empty()) { out += ",\"datatype\":"; jsonStr(out, dt); }
  else if (dtNum >= 0) { out += ",\"datatype\":" + std::to_string(dtNum); }
}
void jsonObj(std::string& out, const PObj& o);
void jsonVal(std::string& out, const PVal& v) {
  if (v.isObj()) { jsonObj(out, *v.obj); return; }
  out += "{\"type\":\"value\",\"value\":"; jsonPrimValue(out, v.prim); jsonDatatype(out, v.prim.datatype, v.prim.datatypeNum); out += "}";
}
void jsonObj(std::string& out, const PObj& o) {
  out += "{\"objectId\":"; jsonStr(out, o.objectId); out += ",\"type\":";
  if (o.typeNull) out += "null"; else jsonStr(out, o.type);
  if (o.isList()) {
    out += ",\"edits\":[";
    bool first = true;
    for (auto& e : o.edits) {
      if (!first) out += ","; first = false;
      switch (e.action) {
        case Edit::INSERT: out += "{\"action\":\"insert\",\"index\":" + std::to_string(e.index) + ",\"elemId\":"; jsonStr(out, e.elemId);
          out += ",\"opId\":"; jsonStr(out, e.opId); out += ",\"value\":"; jsonVal(out, e.value); out += "}"; break;
        case Edit::MULTI_INSERT: out += "{\"action\":\"multi-insert\",\"index\":" + std::to_string(e.index) + ",\"elemId\":"; jsonStr(out, e.elemId);
          if (e.hasDatatype) jsonDatatype(out, e.datatype, e.datatypeNum);
          out += ",\"values\":["; for (size_t i = 0; i < e.values.size(); i++) { if (i) out += ","; jsonPrimValue(out, e.values[i]); } out += "]}"; break;
        case Edit::UPDATE: out += "{\"action\":\"update\",\"index\":" + std::to_string(e.index) + ",\"opId\":"; jsonStr(out, e.opId);
          out += ",\"value\":"; jsonVal(out, e.value); out += "}"; break;
        case Edit::REMOVE: out += "{\"action\":\"remove\",\"index\":" + std::to_string(e.index) + ",\"count\":" + std::to_string(e.count) + "}"; break;
      }
    }
    out += "]}";
  } else {
    out += ",\"props\":{";
    bool first = true;
    for (auto& kv : o.props.items) {
      if (!first) out += ","; first = false;
      jsonStr(out, kv.first); out += ":{";
      bool f2 = true;
      for (auto& ov : kv.second.items) { if (!f2) out += ","; f2 = false; jsonStr(out, ov.first); out += ":"; jsonVal(out, ov.second); }
      out += "}";
    }
    out += "}}";
  }
}
std::string jsonPatch(const PatchResult& r) {
  std::string out = "{\"maxOp\":" + std::to_string(r.maxOp) + ",\"clock\":{";
  bool first = true; for (auto& kv : r.clock) { if (!first) out += ","; first = false; jsonStr(out, kv.first); out += ":" + std::to_string(kv.second); }
  out += "},\"deps\":["; first = true; for (auto& d : r.deps) { if (!first) out += ","; first = false; jsonStr(out, d); }
  out += "],\"pendingChanges\":" + std::to_string(r.pendingChanges);
  if (r.hasActorSeq) { out += ",\"actor\":"; jsonStr(out, r.actor); out += ",\"seq\":" + std::to_string(r.seq); }
  out += ",\"diffs\":"; jsonObj(out, *r.diffs); out += "}";
  return out;
}
char* dupStr(const std::string& s) { char* p = (char*)malloc(s.size() + 1); memcpy(p, s.data(), s.size()); p[s.size()] = 0; return p; }
void setErr(char** err, const char* kind, const std::exception& e) { if (err) *err = dupStr(std::string(kind) + ": " + e.what()); }
std::string jsonMeta(int64_t v) { return v == UNDEF ? "\"undefined\"" : (v == NUL ? "null" : std::to_string(v)); }

#define GUARD(...) try { __VA_ARGS__ } catch (RangeError& e) { setErr(err, "RangeError", e); return 1; } \
  catch (TypeError& e) { setErr(err, "TypeError", e); return 2; } catch (std::exception& e) { setErr(err, "Error", e); return 3; }

std::vector<std::string> splitBlob(const uint8_t* blob, const uint64_t* offsets, size_t n) {
  std::vector<std::string> v; v.reserve(n);
  for (size_t i = 0; i < n; i++) v.emplace_back((const char*)blob + offsets[i], offsets[i + 1] - offsets[i]);
  return v;
}
std::string jsonStrList(const std::vector<std::string>& v) {
  std::string out = "["; for (size_t i = 0; i < v.size(); i++) { if (i) out += ","; jsonStr(out, v[i]); } out += "]"; return out;
}
// packs a list of byte strings as: u64 count, u64 offsets[count+1], bytes
void packBuffers(const std::vector<std::string>& v, uint8_t** out, size_t* len) {
  size_t total = 0; for (auto& s : v) total += s.size();
  size_t hdr = 8 * (v.size() + 2); uint8_t* p = (uint8_t*)malloc(hdr + total); uint64_t* h = (uint64_t*)p;
  h[0] = v.size(); uint64_t off = 0; for (size_t i = 0; i < v.size(); i++) { h[1 + i] = off; memcpy(p + hdr + off, v[i].data(), v[i].size()); off += v[i].size(); }
  h[1 + v.size()] = off; *out = p; *len = hdr + total;
}
}  // namespace

extern "C" {

void orc_free_mem(void* p) { free(p); }
void* orc_new() { return new BackendDoc(); }
void orc_free(void* d) { delete (BackendDoc*)d; }
void* orc_clone(void* d) { return new BackendDoc(*(BackendDoc*)d); }
int orc_load(const uint8_t* buf, size_t len, void** out, char** err) {
  GUARD(*out = new BackendDoc(std::string((const char*)buf, len)); return 0;)
}
// changes are given as one blob + n+1 offsets; out_json may be NULL (skip serialisation, for timing)
int orc_apply_changes(void* d, const uint8_t* blob, const uint64_t* offsets, size_t n, int is_local, char** out_json, char** err) {
  GUARD(PatchResult r = ((BackendDoc*)d)->applyChanges(splitBlob(blob, offsets, n), is_local != 0);
        if (out_json) *out_json = dupStr(jsonPatch(r)); return 0;)
}
int orc_get_patch(void* d, char** out_json, char** err) {
  GUARD(PatchResult r = ((BackendDoc*)d)->getPatch(); if (out_json) *out_json = dupStr(jsonPatch(r)); return 0;)
}
int orc_save(void* d, uint8_t** out, size_t* len, char** err) {
  GUARD(std::string s = ((BackendDoc*)d)->save(); *out = (uint8_t*)malloc(s.size() + 1); memcpy(*out, s.data(), s.size()); *len = s.size(); return 0;)
}
int orc_get_heads(void* d, char** out_json, char** err) { GUARD(*out_json = dupStr(jsonStrList(((BackendDoc*)d)->heads)); return 0;) }
int orc_get_changes(void* d, const char* have_deps_hex /* concatenated 64-char hashes */, uint8_t** out, size_t* len, char** err) {
  GUARD(std::vector<std::string> deps; std::string s(have_deps_hex ? have_deps_hex : "");
        for (size_t i = 0; i + 64 <= s.size(); i += 64) deps.push_back(s.substr(i, 64));
        packBuffers(((BackendDoc*)d)->getChanges(deps), out, len); return 0;)
}
int orc_get_changes_added(void* d_new, void* d_old, uint8_t** out, size_t* len, char** err) {
  GUARD(packBuffers(((BackendDoc*)d_new)->getChangesAdded(*(BackendDoc*)d_old), out, len); return 0;)
}
int orc_get_missing_deps(void* d, const char* heads_hex, char** out_json, char** err) {
  GUARD(std::vector<std::string> hs; std::string s(heads_hex ? heads_hex : "");
        for (size_t i = 0; i + 64 <= s.size(); i += 64) hs.push_back(s.substr(i, 64));
        *out_json = dupStr(jsonStrList(((BackendDoc*)d)->getMissingDeps(hs))); return 0;)
}

int orc_clock_json(void* d, char** out_json, char** err) {
  GUARD(BackendDoc* doc = (BackendDoc*)d; std::string out = "{"; bool first = true;
        for (auto& kv : doc->clock) { if (!first) out += ","; first = false; jsonStr(out, kv.first); out += ":" + std::to_string(kv.second); }
        out += "}"; *out_json = dupStr(out); return 0;)
}
// backend.js:34-45 hashByActor: returns "" when unknown
int orc_hash_by_actor(void* d, const char* actor_hex, int64_t index, char** out, char** err) {
  GUARD(BackendDoc* doc = (BackendDoc*)d; doc->requireHashGraph(); auto it = doc->hashesByActor.find(actor_hex);
        std::string h; if (it != doc->hashesByActor.end() && index >= 0 && index < (int64_t)it->second.size()) h = it->second[index];
        *out = dupStr(h); return 0;)
}
// new.js:1999-2002; *len = (size_t)-1 when the hash is unknown (undefined)
int orc_get_change_by_hash(void* d, const char* hash_hex, uint8_t** out, size_t* len, char** err) {
  GUARD(BackendDoc* doc = (BackendDoc*)d; doc->requireHashGraph(); auto it = doc->changeIndexByHash.find(hash_hex);
        if (it == doc->changeIndexByHash.end() || it->second < 0 || it->second >= (int64_t)doc->changes.size()) { *out = nullptr; *len = (size_t)-1; return 0; }
        const std::string& c = doc->changes[it->second]; *out = (uint8_t*)malloc(c.size() + 1); memcpy(*out, c.data(), c.size()); *len = c.size(); return 0;)
}
int orc_max_op(void* d, int64_t* out) { *out = ((BackendDoc*)d)->maxOp; return 0; }

// Blocks: encoded doc columns + metadata, as the reference's tests inspect them (checkColumns,
// test/new_backend_test.js:7-22)
int orc_blocks_json(void* d, char** out_json, char** err) {
  GUARD(BackendDoc* doc = (BackendDoc*)d; std::string out = "[";
    for (size_t b = 0; b < doc->blocks.size(); b++) {
      const Block& blk = *doc->blocks[b]; if (b) out += ",";
      out += "{\"columns\":{";
      std::vector<Column> cols = encodeDocOps(blk.ops, doc->extraColumnIds);
      for (size_t i = 0; i < cols.size(); i++) { if (i) out += ","; out += "\"" + std::to_string(cols[i].columnId) + "\":\"" + toHex(cols[i].buffer) + "\""; }
      out += "},\"numOps\":" + std::to_string(blk.numOps) + ",\"lastKey\":"; if (blk.hasLastKey) jsonStr(out, blk.lastKey); else out += "\"\\u0000undefined\"";
      out += ",\"numVisible\":" + jsonMeta(blk.numVisible) + ",\"lastObjectActor\":" + jsonMeta(blk.lastObjectActor) + ",\"lastObjectCtr\":" + jsonMeta(blk.lastObjectCtr);
      out += ",\"firstVisibleActor\":" + jsonMeta(blk.firstVisibleActor) + ",\"firstVisibleCtr\":" + jsonMeta(blk.firstVisibleCtr);
      out += ",\"lastVisibleActor\":" + jsonMeta(blk.lastVisibleActor) + ",\"lastVisibleCtr\":" + jsonMeta(blk.lastVisibleCtr);
      out += ",\"bloom\":\"" + toHex(blk.bloom, BLOOM_FILTER_SIZE) + "\"}";
    }
    out += "]"; *out_json = dupStr(out); return 0;)
}

// Doc-ordered op table dump for GPU parity: for every row: idCtr, idActor (doc actor index), succNum;
// succ entries as (ctr, actor) pairs. Arrays are malloc'ed int64.
int orc_dump_ops(void* d, int64_t** rows /* n x 12 */, size_t* n, int64_t** succ /* m x 2 */, size_t* m, char** actors_json, char** err) {
  GUARD(BackendDoc* doc = (BackendDoc*)d; size_t total = 0, st = 0;
    for (auto& b : doc->blocks) { total += b->ops.size(); for (auto& op : b->ops) st += op.succCtr.size(); }
    int64_t* r = (int64_t*)malloc(sizeof(int64_t) * 12 * (total + 1)); int64_t* s = (int64_t*)malloc(sizeof(int64_t) * 2 * (st + 1));
    size_t i = 0, j = 0;
    for (auto& b : doc->blocks) for (auto& op : b->ops) {
      int64_t* o = r + 12 * i++;
      o[0] = op.objCtr; o[1] = op.objActor; o[2] = op.keyCtr; o[3] = op.keyActor; o[4] = op.idCtr; o[5] = op.idActor;
      o[6] = op.insert; o[7] = op.action; o[8] = op.valLen; o[9] = op.succNum(); o[10] = op.chldCtr; o[11] = op.chldActor;
      for (size_t k = 0; k < op.succCtr.size(); k++) { s[2 * j] = op.succCtr[k]; s[2 * j + 1] = op.succActor[k]; j++; }
    }
    *rows = r; *n = total; *succ = s; *m = st; if (actors_json) *actors_json = dupStr(jsonStrList(doc->actorIds)); return 0;)
}

// ---- codec-level entry points (pinned on test/encoding_test.js vectors)
// kind: 0 rle-uint, 1 rle-int, 2 rle-utf8, 3 delta, 4 boolean
int orc_decode_column(int kind, const uint8_t* buf, size_t len, char** out_json, char** err) {
  GUARD(std::string out = "[", b((const char*)buf, len); bool first = true;
    auto put = [&](const RV& v) { if (!first) out += ","; first = false; if (v.null) out += "null"; else if (v.isStr) jsonStr(out, v.str); else out += std::to_string(v.num); };
    if (kind == 4) { BooleanDecoder d(b); while (!d.done()) { if (!first) out += ","; first = false; out += d.readValue() ? "true" : "false"; } }
    else if (kind == 3) { DeltaDecoder d(b); while (!d.done()) put(d.readValue()); }
    else { RLEDecoder d(kind == 0 ? T_UINT : kind == 1 ? T_INT : T_UTF8, b); while (!d.done()) put(d.readValue()); }
    out += "]"; *out_json = dupStr(out); return 0;)
}
// values: int64 array with null mask; for utf8 (kind 2) strs is a blob with n+1 offsets
int orc_encode_column(int kind, const int64_t* vals, const uint8_t* is_null, const uint8_t* strs, const uint64_t* str_off, size_t n,
                      uint8_t** out, size_t* out_len, char** err) {
  GUARD(std::string res;
    if (kind == 4) { BooleanEncoder e; for (size_t i = 0; i < n; i++) e.appendValue(vals[i] != 0); res = e.finish(); }
    else if (kind == 3) { DeltaEncoder e; for (size_t i = 0; i < n; i++) e.appendValue(is_null[i] ? RV() : RV::Num(vals[i])); res = e.finish(); }
    else { RLEEncoder e(kind == 0 ? T_UINT : kind == 1 ? T_INT : T_UTF8);
      for (size_t i = 0; i < n; i++) {
        if (is_null[i]) e.appendValue(RV());
        else if (kind == 2) e.appendValue(RV::Str(std::string((const char*)strs + str_off[i], str_off[i + 1] - str_off[i])));
        else e.appendValue(RV::Num(vals[i]));
      }
      res = e.finish(); }
    *out = (uint8_t*)malloc(res.size() + 1); memcpy(*out, res.data(), res.size()); *out_len = res.size(); return 0;)
}
// kind: 0 uint53, 1 int53, 2 uint32(as uint64 w/ 32-bit check), 3 uint64 (returned as two halves not needed) ; returns value and bytes consumed
int orc_leb_decode(int kind, const uint8_t* buf, size_t len, int64_t* value, size_t* consumed, char** err) {
  GUARD(Decoder d(buf, len); if (kind == 0) *value = d.readUint53(); else if (kind == 1) *value = d.readInt53(); else *value = (int64_t)d.readUint32();
        *consumed = d.offset; return 0;)
}
int orc_leb_encode(int kind, int64_t value, uint8_t* out16, size_t* out_len, char** err) {
  GUARD(Encoder e; if (kind == 0) e.appendUint53(value); else e.appendInt53(value); memcpy(out16, e.buf.data(), e.buf.size()); *out_len = e.buf.size(); return 0;)
}
int orc_sha256(const uint8_t* buf, size_t len, uint8_t out32[32]) { Sha256 s; s.update(buf, len); s.digest(out32); return 0; }

// Decodes one binary change into JSON: header fields + ops as raw column rows (for tests).
int orc_decode_change(const uint8_t* buf, size_t len, char** out_json, char** err) {
  GUARD(DecodedChange ch = decodeChangeColumns(std::string((const char*)buf, len));
    std::vector<Op> ops = readAllOps(ch.columns, CHANGE_COLUMN_IDS, 16, 7, nullptr);
    std::string out = "{\"actor\":"; jsonStr(out, ch.actor); out += ",\"seq\":" + std::to_string(ch.seq) + ",\"startOp\":" + std::to_string(ch.startOp) +
      ",\"time\":" + std::to_string(ch.time) + ",\"message\":"; jsonStr(out, ch.message); out += ",\"hash\":"; jsonStr(out, ch.hash);
    out += ",\"deps\":" + jsonStrList(ch.deps) + ",\"actorIds\":" + jsonStrList(ch.actorIds) + ",\"extraBytes\":\"" + toHex(ch.extraBytes) + "\",\"ops\":[";
    auto nn = [](int64_t v) { return v == NUL ? std::string("null") : std::to_string(v); };
    for (size_t i = 0; i < ops.size(); i++) {
      const Op& o = ops[i]; if (i) out += ",";
      out += "{\"objActor\":" + nn(o.objActor) + ",\"objCtr\":" + nn(o.objCtr) + ",\"keyActor\":" + nn(o.keyActor) + ",\"keyCtr\":" + nn(o.keyCtr) + ",\"keyStr\":";
      if (o.hasKeyStr) jsonStr(out, o.keyStr); else out += "null";
      out += std::string(",\"insert\":") + (o.insert ? "true" : "false") + ",\"action\":" + nn(o.action) + ",\"valLen\":" + nn(o.valLen) + ",\"valRaw\":\"" + toHex(o.valRaw) + "\"";
      out += ",\"chldActor\":" + nn(o.chldActor) + ",\"chldCtr\":" + nn(o.chldCtr) + ",\"pred\":[";
      for (size_t k = 0; k < o.succCtr.size(); k++) { if (k) out += ","; out += "[" + nn(o.succCtr[k]) + "," + nn(o.succActor[k]) + "]"; }
      out += "]}";
    }
    out += "]}"; *out_json = dupStr(out); return 0;)
}
}  // extern "C"
