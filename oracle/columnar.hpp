// ORACLE — TEST INFRASTRUCTURE ONLY (see codec.hpp header).
// CPU restatement of the columnar change / document container format.
//
// Follows (reference paths relative to /root/reference):
//   backend/columnar.js:24-94     magic bytes, chunk types, column / value type tags, column specs
//   backend/columnar.js:300-329   decodeValue
//   backend/columnar.js:525-575   encoderByColumnId / decoderByColumnId / makeDecoders
//   backend/columnar.js:609-652   decodeColumnInfo, encodeColumnInfo, decodeChangeHeader
//   backend/columnar.js:659-708   encodeContainer / decodeContainerHeader (SHA-256 checksum)
//   backend/columnar.js:741-765   decodeChangeColumns
//   backend/columnar.js:798-823   deflateChange / inflateChange   (pako 2.0.3 -> system zlib, raw deflate)
//   backend/columnar.js:983-1038  encodeDocumentHeader / decodeDocumentHeader
//   backend/columnar.js:1052-1067 deflateColumn / inflateColumn
#pragma once
#include <zlib.h>
#include <algorithm>
#include <map>
#include "codec.hpp"

namespace orc {

static const uint8_t MAGIC_BYTES[4] = {0x85, 0x6f, 0x4a, 0x83};
enum { CHUNK_TYPE_DOCUMENT = 0, CHUNK_TYPE_CHANGE = 1, CHUNK_TYPE_DEFLATE = 2 };
static const size_t DEFLATE_MIN_SIZE = 256;
enum ColumnType { GROUP_CARD = 0, ACTOR_ID = 1, INT_RLE = 2, INT_DELTA = 3, BOOLEAN = 4, STRING_RLE = 5, VALUE_LEN = 6, VALUE_RAW = 7 };
static const int COLUMN_TYPE_DEFLATE = 8;
enum ValueType { VT_NULL = 0, VT_FALSE = 1, VT_TRUE = 2, VT_LEB128_UINT = 3, VT_LEB128_INT = 4, VT_IEEE754 = 5,
                 VT_UTF8 = 6, VT_BYTES = 7, VT_COUNTER = 8, VT_TIMESTAMP = 9, VT_MIN_UNKNOWN = 10, VT_MAX_UNKNOWN = 15 };
enum Action { A_MAKE_MAP = 0, A_SET = 1, A_MAKE_LIST = 2, A_DEL = 3, A_MAKE_TEXT = 4, A_INC = 5, A_MAKE_TABLE = 6, A_LINK = 7, NUM_ACTIONS = 8 };

// columnar.js:56-94 column ids
enum {
  COL_OBJ_ACTOR = 0x01, COL_OBJ_CTR = 0x02, COL_KEY_ACTOR = 0x11, COL_KEY_CTR = 0x13, COL_KEY_STR = 0x15,
  COL_ID_ACTOR = 0x21, COL_ID_CTR = 0x23, COL_INSERT = 0x34, COL_ACTION = 0x42, COL_VAL_LEN = 0x56,
  COL_VAL_RAW = 0x57, COL_CHLD_ACTOR = 0x61, COL_CHLD_CTR = 0x63, COL_PRED_NUM = 0x70, COL_PRED_ACTOR = 0x71,
  COL_PRED_CTR = 0x73, COL_SUCC_NUM = 0x80, COL_SUCC_ACTOR = 0x81, COL_SUCC_CTR = 0x83,
  // document change-metadata columns
  DCOL_ACTOR = 0x01, DCOL_SEQ = 0x03, DCOL_MAX_OP = 0x13, DCOL_TIME = 0x23, DCOL_MESSAGE = 0x35,
  DCOL_DEPS_NUM = 0x40, DCOL_DEPS_INDEX = 0x43, DCOL_EXTRA_LEN = 0x56, DCOL_EXTRA_RAW = 0x57
};
static const int DOC_OPS_COLUMN_IDS[] = {0x01, 0x02, 0x11, 0x13, 0x15, 0x21, 0x23, 0x34, 0x42, 0x56, 0x57, 0x61, 0x63, 0x80, 0x81, 0x83};
static const int CHANGE_COLUMN_IDS[] = {0x01, 0x02, 0x11, 0x13, 0x15, 0x21, 0x23, 0x34, 0x42, 0x56, 0x57, 0x61, 0x63, 0x70, 0x71, 0x73};
static const int DOCUMENT_COLUMN_IDS[] = {0x01, 0x03, 0x13, 0x23, 0x35, 0x40, 0x43, 0x56, 0x57};

struct Column { int columnId; std::string buffer; };

inline std::string inflateRaw(const std::string& in) {
  z_stream zs; memset(&zs, 0, sizeof(zs));
  if (inflateInit2(&zs, -15) != Z_OK) throw RangeError("inflateInit failed");
  std::string out; out.resize(std::max<size_t>(in.size() * 4, 1024));
  zs.next_in = (Bytef*)in.data(); zs.avail_in = (uInt)in.size(); size_t produced = 0;
  while (true) {
    zs.next_out = (Bytef*)out.data() + produced; zs.avail_out = (uInt)(out.size() - produced);
    int rc = inflate(&zs, Z_NO_FLUSH);
    produced = out.size() - zs.avail_out;
    if (rc == Z_STREAM_END) break;
    if (rc != Z_OK && rc != Z_BUF_ERROR) { inflateEnd(&zs); throw RangeError("invalid deflate data"); }
    if (zs.avail_out == 0) out.resize(out.size() * 2);
    else if (zs.avail_in == 0) { inflateEnd(&zs); throw RangeError("unexpected end of deflate data"); }
  }
  inflateEnd(&zs); out.resize(produced); return out;
}
inline std::string deflateRaw(const std::string& in) {
  z_stream zs; memset(&zs, 0, sizeof(zs));
  if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw RangeError("deflateInit failed");
  std::string out; out.resize(deflateBound(&zs, in.size()));
  zs.next_in = (Bytef*)in.data(); zs.avail_in = (uInt)in.size();
  zs.next_out = (Bytef*)out.data(); zs.avail_out = (uInt)out.size();
  int rc = deflate(&zs, Z_FINISH); if (rc != Z_STREAM_END) { deflateEnd(&zs); throw RangeError("deflate failed"); }
  out.resize(zs.total_out); deflateEnd(&zs); return out;
}

// columnar.js:688-708
struct ContainerHeader { int chunkType; int64_t chunkLength; std::string chunkData; std::string hash; /* hex */ };
inline ContainerHeader decodeContainerHeader(Decoder& d, bool computeHash) {
  std::string magic = d.readRawBytes(4);
  if (memcmp(magic.data(), MAGIC_BYTES, 4) != 0) throw RangeError("Data does not begin with magic bytes 85 6f 4a 83");
  std::string expected = d.readRawBytes(4);
  size_t hashStart = d.offset;
  ContainerHeader h; h.chunkType = d.readByte(); h.chunkLength = d.readUint53();
  h.chunkData = d.readRawBytes((size_t)h.chunkLength);
  if (computeHash) {
    Sha256 s; s.update(d.buf + hashStart, d.offset - hashStart); uint8_t dig[32]; s.digest(dig);
    if (memcmp(dig, expected.data(), 4) != 0) throw RangeError("checksum does not match data");
    h.hash = toHex(dig, 32);
  }
  return h;
}
// columnar.js:659-686
inline std::string encodeContainer(int chunkType, const std::string& body, std::string* hashHex = nullptr) {
  Encoder header; header.appendByte((uint8_t)chunkType); header.appendUint53((int64_t)body.size());
  Sha256 s; s.update((const uint8_t*)header.buf.data(), header.buf.size()); s.update((const uint8_t*)body.data(), body.size());
  uint8_t dig[32]; s.digest(dig);
  if (hashHex) *hashHex = toHex(dig, 32);
  std::string out((const char*)MAGIC_BYTES, 4); out.append((const char*)dig, 4); out += header.buf; out += body; return out;
}
// columnar.js:609-624
inline std::vector<Column> decodeColumnInfo(Decoder& d, std::vector<int64_t>& lens) {
  const uint32_t MASK = ~(uint32_t)COLUMN_TYPE_DEFLATE;
  int64_t lastColumnId = -1; std::vector<Column> cols; int64_t n = d.readUint53();
  for (int64_t i = 0; i < n; i++) {
    int64_t columnId = d.readUint53(), bufferLen = d.readUint53();
    if (lastColumnId >= 0 && ((uint32_t)columnId & MASK) <= ((uint32_t)lastColumnId & MASK)) throw RangeError("Columns must be in ascending order");
    lastColumnId = columnId; cols.push_back({(int)columnId, std::string()}); lens.push_back(bufferLen);
  }
  return cols;
}
// columnar.js:626-633
inline void encodeColumnInfo(Encoder& e, const std::vector<Column>& cols) {
  int64_t n = 0; for (auto& c : cols) if (!c.buffer.empty()) n++;
  e.appendUint53(n);
  for (auto& c : cols) if (!c.buffer.empty()) { e.appendUint53(c.columnId); e.appendUint53((int64_t)c.buffer.size()); }
}

// columnar.js:813-823 inflateChange
inline std::string inflateChange(const std::string& buffer) {
  Decoder d(buffer); ContainerHeader h = decodeContainerHeader(d, false);
  if (h.chunkType != CHUNK_TYPE_DEFLATE) throw RangeError("Unexpected chunk type: " + std::to_string(h.chunkType));
  std::string raw = inflateRaw(h.chunkData);
  Encoder e; e.appendRaw(buffer.substr(0, 8)); e.appendByte(CHUNK_TYPE_CHANGE); e.appendUint53((int64_t)raw.size()); e.appendRaw(raw);
  return e.buf;
}
// columnar.js:798-808 deflateChange
inline std::string deflateChange(const std::string& buffer) {
  Decoder d(buffer); ContainerHeader h = decodeContainerHeader(d, false);
  if (h.chunkType != CHUNK_TYPE_CHANGE) throw RangeError("Unexpected chunk type: " + std::to_string(h.chunkType));
  std::string comp = deflateRaw(h.chunkData);
  Encoder e; e.appendRaw(buffer.substr(0, 8)); e.appendByte(CHUNK_TYPE_DEFLATE); e.appendUint53((int64_t)comp.size()); e.appendRaw(comp);
  return e.buf;
}

// decoded change (columnar.js:741-765 decodeChangeColumns)
struct DecodedChange {
  std::string actor; int64_t seq = 0, startOp = 0, time = 0, maxOp = 0; std::string message;
  std::vector<std::string> deps;      // hex
  std::vector<std::string> actorIds;  // hex; [0] = author
  std::string hash;                   // hex
  std::vector<Column> columns; std::string extraBytes; bool hasExtra = false;
  std::string buffer;                 // original (possibly deflated) bytes
};
inline DecodedChange decodeChangeColumns(const std::string& input) {
  std::string inflated; const std::string* bufp = &input;
  if (input.size() > 8 && (uint8_t)input[8] == CHUNK_TYPE_DEFLATE) { inflated = inflateChange(input); bufp = &inflated; }
  Decoder d(*bufp); ContainerHeader h = decodeContainerHeader(d, true);
  if (!d.done()) throw RangeError("Encoded change has trailing data");
  if (h.chunkType != CHUNK_TYPE_CHANGE) throw RangeError("Unexpected chunk type: " + std::to_string(h.chunkType));
  Decoder c(h.chunkData); DecodedChange ch;
  // columnar.js:635-652 decodeChangeHeader
  int64_t numDeps = c.readUint53();
  for (int64_t i = 0; i < numDeps; i++) ch.deps.push_back(toHex(c.readRawBytes(32)));
  ch.actor = c.readHexString(); ch.seq = c.readUint53(); ch.startOp = c.readUint53(); ch.time = c.readInt53();
  ch.message = c.readPrefixedBytes(); ch.actorIds.push_back(ch.actor);
  int64_t numActorIds = c.readUint53();
  for (int64_t i = 0; i < numActorIds; i++) ch.actorIds.push_back(c.readHexString());
  std::vector<int64_t> lens; ch.columns = decodeColumnInfo(c, lens);
  for (size_t i = 0; i < ch.columns.size(); i++) {
    if (ch.columns[i].columnId & COLUMN_TYPE_DEFLATE) throw RangeError("change must not contain deflated columns");
    ch.columns[i].buffer = c.readRawBytes((size_t)lens[i]);
  }
  if (!c.done()) { ch.extraBytes = c.readRawBytes(c.len - c.offset); ch.hasExtra = true; }
  ch.hash = h.hash; ch.buffer = input;
  return ch;
}

// columnar.js:1006-1038
struct DocHeader {
  std::vector<Column> changesColumns, opsColumns; std::vector<std::string> actorIds, heads; std::vector<int64_t> headsIndexes; std::string extraBytes;
};
inline void inflateColumn(Column& c) {
  if (c.columnId & COLUMN_TYPE_DEFLATE) { c.buffer = inflateRaw(c.buffer); c.columnId ^= COLUMN_TYPE_DEFLATE; }
}
inline void deflateColumn(Column& c) {
  if (c.buffer.size() >= DEFLATE_MIN_SIZE) { c.buffer = deflateRaw(c.buffer); c.columnId |= COLUMN_TYPE_DEFLATE; }
}
inline DocHeader decodeDocumentHeader(const std::string& buffer) {
  Decoder dd(buffer); ContainerHeader h = decodeContainerHeader(dd, true);
  if (!dd.done()) throw RangeError("Encoded document has trailing data");
  if (h.chunkType != CHUNK_TYPE_DOCUMENT) throw RangeError("Unexpected chunk type: " + std::to_string(h.chunkType));
  Decoder d(h.chunkData); DocHeader doc;
  int64_t numActors = d.readUint53(); for (int64_t i = 0; i < numActors; i++) doc.actorIds.push_back(d.readHexString());
  int64_t numHeads = d.readUint53(); for (int64_t i = 0; i < numHeads; i++) doc.heads.push_back(toHex(d.readRawBytes(32)));
  std::vector<int64_t> l1, l2; doc.changesColumns = decodeColumnInfo(d, l1); doc.opsColumns = decodeColumnInfo(d, l2);
  for (size_t i = 0; i < doc.changesColumns.size(); i++) { doc.changesColumns[i].buffer = d.readRawBytes((size_t)l1[i]); inflateColumn(doc.changesColumns[i]); }
  for (size_t i = 0; i < doc.opsColumns.size(); i++) { doc.opsColumns[i].buffer = d.readRawBytes((size_t)l2[i]); inflateColumn(doc.opsColumns[i]); }
  if (!d.done()) for (int64_t i = 0; i < numHeads; i++) doc.headsIndexes.push_back(d.readUint53());
  doc.extraBytes = d.readRawBytes(d.len - d.offset);
  return doc;
}
// columnar.js:983-1004
inline std::string encodeDocumentHeader(DocHeader doc) {
  for (auto& c : doc.changesColumns) deflateColumn(c);
  for (auto& c : doc.opsColumns) deflateColumn(c);
  Encoder e; e.appendUint53((int64_t)doc.actorIds.size()); for (auto& a : doc.actorIds) e.appendHexString(a);
  std::sort(doc.heads.begin(), doc.heads.end());
  e.appendUint53((int64_t)doc.heads.size()); for (auto& hd : doc.heads) e.appendRaw(fromHex(hd));
  encodeColumnInfo(e, doc.changesColumns); encodeColumnInfo(e, doc.opsColumns);
  for (auto& c : doc.changesColumns) e.appendRaw(c.buffer);
  for (auto& c : doc.opsColumns) e.appendRaw(c.buffer);
  for (auto ix : doc.headsIndexes) e.appendUint53(ix);
  e.appendRaw(doc.extraBytes);
  return encodeContainer(CHUNK_TYPE_DOCUMENT, e.buf);
}

// A primitive JS value as produced by decodeValue (columnar.js:300-329)
struct Prim {
  enum K { P_NULL, P_BOOL, P_INT, P_FLOAT, P_STR, P_BYTES } k = P_NULL;
  bool b = false; int64_t i = 0; double f = 0; std::string s;
  std::string datatype; int datatypeNum = -1;   // 'uint','int','float64','counter','timestamp' or unknown tag number
  // JS typeof
  int typeOf() const { switch (k) { case P_BOOL: return 1; case P_INT: case P_FLOAT: return 2; case P_STR: return 3; default: return 0; } }
  bool sameDatatype(const Prim& o) const { return datatype == o.datatype && datatypeNum == o.datatypeNum; }
  double num() const { return k == P_INT ? (double)i : f; }
};
inline Prim decodeValue(int64_t sizeTag, const std::string& bytes) {
  Prim p;
  if (sizeTag == VT_NULL) return p;
  if (sizeTag == VT_FALSE) { p.k = Prim::P_BOOL; p.b = false; return p; }
  if (sizeTag == VT_TRUE) { p.k = Prim::P_BOOL; p.b = true; return p; }
  int tag = (int)(sizeTag % 16);
  if (tag == VT_UTF8) { p.k = Prim::P_STR; p.s = bytes; return p; }
  if (tag == VT_LEB128_UINT) { Decoder d(bytes); p.k = Prim::P_INT; p.i = d.readUint53(); p.datatype = "uint"; return p; }
  if (tag == VT_LEB128_INT) { Decoder d(bytes); p.k = Prim::P_INT; p.i = d.readInt53(); p.datatype = "int"; return p; }
  if (tag == VT_IEEE754) {
    if (bytes.size() != 8) throw RangeError("Invalid length for floating point number: " + std::to_string(bytes.size()));
    p.k = Prim::P_FLOAT; memcpy(&p.f, bytes.data(), 8); p.datatype = "float64"; return p;
  }
  if (tag == VT_COUNTER) { Decoder d(bytes); p.k = Prim::P_INT; p.i = d.readInt53(); p.datatype = "counter"; return p; }
  if (tag == VT_TIMESTAMP) { Decoder d(bytes); p.k = Prim::P_INT; p.i = d.readInt53(); p.datatype = "timestamp"; return p; }
  p.k = Prim::P_BYTES; p.s = bytes; p.datatypeNum = tag; return p;
}

}  // namespace orc
