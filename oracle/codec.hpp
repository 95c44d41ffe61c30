// ORACLE — TEST INFRASTRUCTURE ONLY.
// CPU restatement of automerge-classic's byte/varint/run-length codecs and SHA-256.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// use anything under oracle/.  The product (automerge_classic_b200/) never links or calls this.
//
// Follows (reference paths relative to /root/reference):
//   backend/encoding.js:57-286    Encoder (LEB128 append)
//   backend/encoding.js:293-534   Decoder (LEB128 read, range errors)
//   backend/encoding.js:558-783   RLEEncoder        (restated as a batch canonical encoder)
//   backend/encoding.js:789-920   RLEDecoder
//   backend/encoding.js:932-998   DeltaEncoder
//   backend/encoding.js:1004-1051 DeltaDecoder
//   backend/encoding.js:1061-1135 BooleanEncoder
//   backend/encoding.js:1141-1207 BooleanDecoder
//   SHA-256: fast-sha256@1.3.0 is not vendored in the reference tree; this is FIPS 180-4,
//   pinned by the golden checksums in test/columnar_test.js:17 and test/new_backend_test.js:1860.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <algorithm>

namespace orc {

struct RangeError : std::runtime_error { using std::runtime_error::runtime_error; };
struct TypeError : std::runtime_error { using std::runtime_error::runtime_error; };

static const int64_t MAX_SAFE = (1LL << 53) - 1;

// ---------------------------------------------------------------- SHA-256 (FIPS 180-4)
struct Sha256 {
  uint32_t h[8]; uint8_t block[64]; size_t fill = 0; uint64_t total = 0;
  Sha256() {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                                   0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(h, iv, sizeof(iv));
  }
  static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
  void compress(const uint8_t* p) {
    static const uint32_t K[64] = {
      0x428a2f98,0x71374491,0xb5c0fbcf,0xe9b5dba5,0x3956c25b,0x59f111f1,0x923f82a4,0xab1c5ed5,
      0xd807aa98,0x12835b01,0x243185be,0x550c7dc3,0x72be5d74,0x80deb1fe,0x9bdc06a7,0xc19bf174,
      0xe49b69c1,0xefbe4786,0x0fc19dc6,0x240ca1cc,0x2de92c6f,0x4a7484aa,0x5cb0a9dc,0x76f988da,
      0x983e5152,0xa831c66d,0xb00327c8,0xbf597fc7,0xc6e00bf3,0xd5a79147,0x06ca6351,0x14292967,
      0x27b70a85,0x2e1b2138,0x4d2c6dfc,0x53380d13,0x650a7354,0x766a0abb,0x81c2c92e,0x92722c85,
      0xa2bfe8a1,0xa81a664b,0xc24b8b70,0xc76c51a3,0xd192e819,0xd6990624,0xf40e3585,0x106aa070,
      0x19a4c116,0x1e376c08,0x2748774c,0x34b0bcb5,0x391c0cb3,0x4ed8aa4a,0x5b9cca4f,0x682e6ff3,
      0x748f82ee,0x78a5636f,0x84c87814,0x8cc70208,0x90befffa,0xa4506ceb,0xbef9a3f7,0xc67178f2};
    uint32_t w[64];
    for (int i = 0; i < 16; i++)
      w[i] = (uint32_t)p[4*i] << 24 | (uint32_t)p[4*i+1] << 16 | (uint32_t)p[4*i+2] << 8 | p[4*i+3];
    for (int i = 16; i < 64; i++) {
      uint32_t s0 = rotr(w[i-15], 7) ^ rotr(w[i-15], 18) ^ (w[i-15] >> 3);
      uint32_t s1 = rotr(w[i-2], 17) ^ rotr(w[i-2], 19) ^ (w[i-2] >> 10);
      w[i] = w[i-16] + s0 + w[i-7] + s1;
    }
    uint32_t a=h[0],b=h[1],c=h[2],d=h[3],e=h[4],f=h[5],g=h[6],hh=h[7];
    for (int i = 0; i < 64; i++) {
      uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
      uint32_t ch = (e & f) ^ (~e & g);
      uint32_t t1 = hh + S1 + ch + K[i] + w[i];
      uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
      uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
      uint32_t t2 = S0 + mj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0]+=a; h[1]+=b; h[2]+=c; h[3]+=d; h[4]+=e; h[5]+=f; h[6]+=g; h[7]+=hh;
  }
  void update(const uint8_t* p, size_t n) {
    total += n;
    while (n > 0) {
      size_t take = std::min(n, 64 - fill);
      memcpy(block + fill, p, take); fill += take; p += take; n -= take;
      if (fill == 64) { compress(block); fill = 0; }
    }
  }
  void digest(uint8_t out[32]) {
    uint64_t bits = total * 8;
    uint8_t pad = 0x80; update(&pad, 1);
    uint8_t z = 0; while (fill != 56) update(&z, 1);
    uint8_t len[8]; for (int i = 0; i < 8; i++) len[i] = (uint8_t)(bits >> (56 - 8*i));
    update(len, 8);
    for (int i = 0; i < 8; i++) { out[4*i] = h[i] >> 24; out[4*i+1] = h[i] >> 16; out[4*i+2] = h[i] >> 8; out[4*i+3] = h[i]; }
  }
};

inline std::string toHex(const uint8_t* p, size_t n) {
  static const char* d = "0123456789abcdef"; std::string s; s.reserve(2*n);
  for (size_t i = 0; i < n; i++) { s.push_back(d[p[i] >> 4]); s.push_back(d[p[i] & 15]); }
  return s;
}
inline std::string toHex(const std::string& b) { return toHex((const uint8_t*)b.data(), b.size()); }
// encoding.js:22-35 hexStringToBytes
inline std::string fromHex(const std::string& s) {
  if (s.size() % 2) throw RangeError("value is not hexadecimal");
  std::string out; out.reserve(s.size() / 2);
  auto nib = [](char c) -> int {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    throw RangeError("value is not hexadecimal");
  };
  for (size_t i = 0; i < s.size(); i += 2) out.push_back((char)(nib(s[i]) << 4 | nib(s[i+1])));
  return out;
}

// ---------------------------------------------------------------- Encoder (encoding.js:57-286)
struct Encoder {
  std::string buf;
  void appendByte(uint8_t b) { buf.push_back((char)b); }
  // encoding.js:136-160 appendUint53 (range check) -> LEB128
  void appendUint53(int64_t v) {
    if (v < 0 || v > MAX_SAFE) throw RangeError("number out of range");
    uint64_t u = (uint64_t)v;
    do { uint8_t b = u & 0x7f; u >>= 7; if (u) b |= 0x80; buf.push_back((char)b); } while (u);
  }
  // encoding.js:162-178 appendInt53 -> signed LEB128
  void appendInt53(int64_t v) {
    if (v < -MAX_SAFE || v > MAX_SAFE) throw RangeError("number out of range");
    while (true) {
      uint8_t b = v & 0x7f; v >>= 7;  // arithmetic shift
      if ((v == 0 && !(b & 0x40)) || (v == -1 && (b & 0x40))) { buf.push_back((char)b); return; }
      buf.push_back((char)(b | 0x80));
    }
  }
  void appendRaw(const std::string& s) { buf += s; }
  void appendRaw(const uint8_t* p, size_t n) { buf.append((const char*)p, n); }
  void appendPrefixed(const std::string& s) { appendUint53((int64_t)s.size()); buf += s; }
  void appendHexString(const std::string& hex) { appendPrefixed(fromHex(hex)); }
};

// ---------------------------------------------------------------- Decoder (encoding.js:293-534)
struct Decoder {
  const uint8_t* buf = nullptr; size_t len = 0, offset = 0;
  Decoder() {}
  Decoder(const uint8_t* b, size_t n) : buf(b), len(n) {}
  explicit Decoder(const std::string& s) : buf((const uint8_t*)s.data()), len(s.size()) {}
  bool done() const { return offset == len; }
  void skip(size_t n) { if (offset + n > len) throw RangeError("cannot skip beyond end of buffer"); offset += n; }
  uint8_t readByte() { offset += 1; return offset <= len ? buf[offset - 1] : 0; }
  // encoding.js:416-441 readUint64
  uint64_t readUint64() {
    uint64_t result = 0; int shift = 0;
    while (offset < len) {
      uint8_t b = buf[offset];
      if (shift == 63 && (b & 0xfe) != 0) throw RangeError("number out of range");
      result |= (uint64_t)(b & 0x7f) << shift; shift += 7; offset++;
      if ((b & 0x80) == 0) return result;
    }
    throw RangeError("buffer ended with incomplete number");
  }
  // encoding.js:450-488 readInt64
  int64_t readInt64() {
    uint64_t result = 0; int shift = 0;
    while (offset < len) {
      uint8_t b = buf[offset];
      if (shift == 63 && b != 0 && b != 0x7f) throw RangeError("number out of range");
      result |= (uint64_t)(b & 0x7f) << shift; shift += 7; offset++;
      if ((b & 0x80) == 0) {
        if ((b & 0x40) && shift < 64) result |= ~(uint64_t)0 << shift;
        return (int64_t)result;
      }
    }
    throw RangeError("buffer ended with incomplete number");
  }
  // encoding.js:389-395
  int64_t readUint53() {
    uint64_t v = readUint64();
    if (v > (uint64_t)MAX_SAFE) throw RangeError("number out of range");
    return (int64_t)v;
  }
  // encoding.js:402-408
  int64_t readInt53() {
    int64_t v = readInt64();
    if (v < -MAX_SAFE || v > MAX_SAFE) throw RangeError("number out of range");
    return v;
  }
  // encoding.js:341-354
  uint32_t readUint32() {
    uint64_t v; size_t start = offset;
    try { v = readUint64(); } catch (RangeError& e) { throw; }
    if (v > 0xffffffffULL || offset - start > 5) throw RangeError("number out of range");
    return (uint32_t)v;
  }
  std::string readRawBytes(size_t n) {
    if (offset + n > len) throw RangeError("subarray exceeds buffer size");
    std::string s((const char*)buf + offset, n); offset += n; return s;
  }
  std::string readPrefixedBytes() { return readRawBytes((size_t)readUint53()); }
  std::string readHexString() { return toHex(readPrefixedBytes()); }
};

// A decoded column value: null, number or string.
struct RV {
  bool null = true; int64_t num = 0; std::string str; bool isStr = false;
  static RV Null() { return RV(); }
  static RV Num(int64_t v) { RV r; r.null = false; r.num = v; return r; }
  static RV Str(std::string s) { RV r; r.null = false; r.isStr = true; r.str = std::move(s); return r; }
  bool operator==(const RV& o) const {
    if (null != o.null) return false; if (null) return true;
    if (isStr != o.isStr) return false; return isStr ? str == o.str : num == o.num;
  }
  bool operator!=(const RV& o) const { return !(*this == o); }
};

enum RleType { T_INT, T_UINT, T_UTF8 };

// ---------------------------------------------------------------- RLEDecoder (encoding.js:789-920)
struct RLEDecoder : Decoder {
  RleType type = T_UINT; RV lastValue; bool haveLast = false; int64_t count = 0;
  enum St { S_NONE, S_REP, S_LIT, S_NULLS } state = S_NONE;
  RLEDecoder() {}
  RLEDecoder(RleType t, const uint8_t* b, size_t n) : Decoder(b, n), type(t) {}
  RLEDecoder(RleType t, const std::string& s) : Decoder(s), type(t) {}
  bool done() const { return count == 0 && offset == len; }
  void reset() { offset = 0; haveLast = false; lastValue = RV(); count = 0; state = S_NONE; }
  RV readRawValue() {
    if (type == T_INT) return RV::Num(readInt53());
    if (type == T_UINT) return RV::Num(readUint53());
    return RV::Str(readPrefixedBytes());
  }
  // encoding.js:865-887
  void readRecord() {
    count = readInt53();
    if (count > 1) {
      RV value = readRawValue();
      if ((state == S_REP || state == S_LIT) && haveLast && lastValue == value)
        throw RangeError("Successive repetitions with the same value are not allowed");
      state = S_REP; lastValue = value; haveLast = true;
    } else if (count == 1) {
      throw RangeError("Repetition count of 1 is not allowed, use a literal instead");
    } else if (count < 0) {
      count = -count;
      if (state == S_LIT) throw RangeError("Successive literals are not allowed");
      state = S_LIT;
    } else {
      if (state == S_NULLS) throw RangeError("Successive null runs are not allowed");
      count = readUint53();
      if (count == 0) throw RangeError("Zero-length null runs are not allowed");
      lastValue = RV(); haveLast = true; state = S_NULLS;
    }
  }
  // encoding.js:820-832
  RV readValue() {
    if (done()) return RV();
    if (count == 0) readRecord();
    count -= 1;
    if (state == S_LIT) {
      RV value = readRawValue();
      if (haveLast && value == lastValue) throw RangeError("Repetition of values is not allowed in literal");
      lastValue = value; haveLast = true;
      return value;
    }
    return lastValue;
  }
};

// ---------------------------------------------------------------- DeltaDecoder (encoding.js:1004-1051)
struct DeltaDecoder : RLEDecoder {
  int64_t absoluteValue = 0;
  DeltaDecoder() { type = T_INT; }
  DeltaDecoder(const uint8_t* b, size_t n) : RLEDecoder(T_INT, b, n) {}
  explicit DeltaDecoder(const std::string& s) : RLEDecoder(T_INT, s) {}
  void reset() { RLEDecoder::reset(); absoluteValue = 0; }
  RV readValue() {
    RV v = RLEDecoder::readValue();
    if (v.null) return v;
    absoluteValue += v.num;
    return RV::Num(absoluteValue);
  }
};

// ---------------------------------------------------------------- BooleanDecoder (encoding.js:1141-1207)
struct BooleanDecoder : Decoder {
  bool lastValue = true, firstRun = true; int64_t count = 0;
  BooleanDecoder() {}
  BooleanDecoder(const uint8_t* b, size_t n) : Decoder(b, n) {}
  explicit BooleanDecoder(const std::string& s) : Decoder(s) {}
  bool done() const { return count == 0 && offset == len; }
  void reset() { offset = 0; lastValue = true; firstRun = true; count = 0; }
  bool readValue() {
    if (done()) return false;
    while (count == 0) {
      count = readUint53();
      lastValue = !lastValue;
      if (count == 0 && !firstRun) throw RangeError("Zero-length runs are not allowed");
      firstRun = false;
    }
    count -= 1;
    return lastValue;
  }
};

// ---------------------------------------------------------------- RLEEncoder (encoding.js:558-783)
// The reference keeps an incremental state machine (empty/loneValue/repetition/literal/nulls).
// Its output is the unique canonical form (maximal null runs; maximal runs of >=2 equal values
// as repetitions; everything else gathered into literals; all-null column -> zero bytes), so this
// restatement buffers (value, repetitions) pairs and emits that canonical form in finish().
struct RLEEncoder {
  RleType type; std::vector<std::pair<RV, int64_t>> runs;  // merged runs of equal values
  explicit RLEEncoder(RleType t = T_UINT) : type(t) {}
  virtual ~RLEEncoder() {}
  virtual void appendValue(const RV& v, int64_t repetitions = 1) { rawAppend(v, repetitions); }
  void rawAppend(const RV& v, int64_t repetitions) {
    if (repetitions <= 0) return;
    if (!runs.empty() && runs.back().first == v) runs.back().second += repetitions;
    else runs.emplace_back(v, repetitions);
  }
  void appendRawValue(Encoder& e, const RV& v) const {
    if (type == T_INT) e.appendInt53(v.num);
    else if (type == T_UINT) e.appendUint53(v.num);
    else e.appendPrefixed(v.str);
  }
  std::string finish() const {
    Encoder e; std::vector<const RV*> literal;
    auto flushLiteral = [&]() {
      if (literal.empty()) return;
      e.appendInt53(-(int64_t)literal.size());
      for (auto* v : literal) appendRawValue(e, *v);
      literal.clear();
    };
    // encoding.js:778-782: nothing is written if only nulls have been seen
    bool allNull = true; for (auto& r : runs) if (!r.first.null) allNull = false;
    if (allNull) return std::string();
    for (auto& r : runs) {
      if (r.first.null) { flushLiteral(); e.appendInt53(0); e.appendUint53(r.second); }
      else if (r.second >= 2) { flushLiteral(); e.appendInt53(r.second); appendRawValue(e, r.first); }
      else literal.push_back(&r.first);
    }
    flushLiteral();
    return e.buf;
  }
};

// ---------------------------------------------------------------- DeltaEncoder (encoding.js:932-998)
struct DeltaEncoder : RLEEncoder {
  int64_t absoluteValue = 0;
  DeltaEncoder() : RLEEncoder(T_INT) {}
  void appendValue(const RV& v, int64_t repetitions = 1) override {
    if (repetitions <= 0) return;
    if (!v.null) {
      rawAppend(RV::Num(v.num - absoluteValue), 1);
      absoluteValue = v.num;
      if (repetitions > 1) rawAppend(RV::Num(0), repetitions - 1);
    } else rawAppend(v, repetitions);
  }
};

// ---------------------------------------------------------------- BooleanEncoder (encoding.js:1061-1135)
struct BooleanEncoder {
  Encoder e; bool lastValue = false; int64_t count = 0;
  void appendValue(bool v, int64_t repetitions = 1) {
    if (repetitions <= 0) return;
    if (lastValue == v) count += repetitions;
    else { e.appendUint53(count); lastValue = v; count = repetitions; }
  }
  std::string finish() { if (count > 0) { e.appendUint53(count); count = 0; } return e.buf; }
};

// JavaScript's `<` on strings compares UTF-16 code units (new.js:84, 250, 1159 compare map keys this way). Keys are held
// as UTF-8 here; byte order equals code point order, which differs from code unit order exactly when a supplementary-plane
// character (surrogate pair, units 0xD800-0xDFFF) meets one of U+E000..U+FFFF. Compared unit by unit after decoding.
inline void utf16_units(const std::string& s, std::vector<uint16_t>& out) {
  out.clear(); size_t i = 0; const size_t n = s.size();
  while (i < n) {
    const unsigned char c = (unsigned char)s[i]; uint32_t cp; size_t len;
    if (c < 0x80) { cp = c; len = 1; } else if ((c >> 5) == 6) { cp = c & 0x1f; len = 2; } else if ((c >> 4) == 14) { cp = c & 0x0f; len = 3; } else if ((c >> 3) == 30) { cp = c & 0x07; len = 4; } else { cp = 0xfffd; len = 1; }
    for (size_t k = 1; k < len; k++) cp = (i + k < n) ? ((cp << 6) | ((unsigned char)s[i + k] & 0x3f)) : 0xfffd;
    i += len;
    if (cp >= 0x10000) { cp -= 0x10000; out.push_back((uint16_t)(0xd800 | (cp >> 10))); out.push_back((uint16_t)(0xdc00 | (cp & 0x3ff))); } else out.push_back((uint16_t)cp);
  }
}
inline bool js_less(const std::string& a, const std::string& b) {
  bool ascii = true; for (unsigned char c : a) if (c >= 0xe0) { ascii = false; break; }
  if (ascii) for (unsigned char c : b) if (c >= 0xe0) { ascii = false; break; }
  if (ascii) return a < b;   // no three- or four-byte sequences: byte order is code unit order
  std::vector<uint16_t> x, y; utf16_units(a, x); utf16_units(b, y);
  return std::lexicographical_compare(x.begin(), x.end(), y.begin(), y.end());
}

}  // namespace orc
