// amgpu — kernels #1b: raw DEFLATE (RFC 1951) decoder, one thread per compressed change.
//
// Replaces (reference paths relative to /root/reference):
//   backend/columnar.js:813-823  inflateChange (pako.inflateRaw on the chunk body of a chunk of type 2)
//
// Changes of 256 bytes and more travel DEFLATEd (columnar.js:738); in the text trace those are the
// merge changes (many deps), about 1 % of the batch and ~300 bytes each. One thread decodes one change:
// canonical-Huffman decode bit by bit (code lengths <= 15), literal/length + distance alphabets, stored,
// fixed and dynamic blocks. Two passes over the same stream: pass 0 only counts the output (so the
// inflated changes can be laid out behind the batch with a scan), pass 1 writes it. The work per thread
// is serial but small, and the ~10^4 streams of a batch decode concurrently.
#pragma once
#include "decode.cuh"

namespace amg {

struct BitSource {
  const u8* p; u32 pos, end; u32 buf; int cnt; bool err;
  HD BitSource(const u8* p_, u32 begin, u32 end_) : p(p_), pos(begin), end(end_), buf(0), cnt(0), err(false) {}
  HD u32 bits(int n) {   // n <= 16
    while (cnt < n) { if (pos >= end) { err = true; return 0; } buf |= (u32)p[pos++] << cnt; cnt += 8; }
    const u32 v = buf & ((1u << n) - 1u); buf >>= n; cnt -= n; return v;
  }
  HD void alignByte() { buf = 0; cnt = 0; }
};

// canonical Huffman code: count[len] codes of each length, symbols ordered by (length, symbol value)
template <int MAXSYM> struct HuffCode {
  uint16_t count[16]; uint16_t symbol[MAXSYM];
  HD bool build(const u8* lengths, int n) {   // false: over-subscribed or incomplete set of lengths
    for (int i = 0; i < 16; i++) count[i] = 0;
    for (int i = 0; i < n; i++) count[lengths[i]]++;
    if (count[0] == n) return true;           // no codes at all: legal for the distance alphabet, decode() then fails
    int left = 1;
    for (int len = 1; len < 16; len++) { left <<= 1; left -= count[len]; if (left < 0) return false; }
    uint16_t offs[16]; offs[1] = 0;
    for (int len = 1; len < 15; len++) offs[len + 1] = (uint16_t)(offs[len] + count[len]);
    for (int i = 0; i < n; i++) if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
    return left == 0 || (n - count[0]) == 1;  // complete, or the single-code case RFC 1951 allows
  }
  HD int decode(BitSource& b) const {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len < 16; len++) {
      code |= (int)b.bits(1); if (b.err) return -1;
      const int c = count[len];
      if (code - c < first) return symbol[index + (code - first)];
      index += c; first += c; first <<= 1; code <<= 1;
    }
    return -1;
  }
};

// Decodes src[begin, end) into dst (nullptr = only count). Returns a KErr code (0 = ok); *produced = output length.
HD u32 inflate_raw(const u8* src, u32 begin, u32 end, u8* dst, u32 dstCap, u32* produced) {
  const uint16_t lenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  const u8 lenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  const uint16_t distBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  const u8 distExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  const u8 clOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  BitSource b(src, begin, end); u32 out = 0; *produced = 0;
  HuffCode<288> lit; HuffCode<30> dist; u8 lengths[320];
  while (true) {
    const u32 last = b.bits(1), type = b.bits(2);
    if (b.err) return KE_DEFLATE;
    if (type == 0) {            // stored
      b.alignByte();
      if (b.pos + 4 > b.end) return KE_DEFLATE;
      const u32 len = src[b.pos] | ((u32)src[b.pos + 1] << 8), nlen = src[b.pos + 2] | ((u32)src[b.pos + 3] << 8);
      if ((len ^ 0xffffu) != nlen) return KE_DEFLATE;
      b.pos += 4;
      if (b.pos + len > b.end) return KE_DEFLATE;
      if (dst) { if (out + len > dstCap) return KE_DEFLATE; for (u32 i = 0; i < len; i++) dst[out + i] = src[b.pos + i]; }
      out += len; b.pos += len;
    } else if (type == 1 || type == 2) {
      if (type == 1) {          // fixed code (RFC 1951 3.2.6)
        for (int i = 0; i < 144; i++) lengths[i] = 8;
        for (int i = 144; i < 256; i++) lengths[i] = 9;
        for (int i = 256; i < 280; i++) lengths[i] = 7;
        for (int i = 280; i < 288; i++) lengths[i] = 8;
        lit.build(lengths, 288);
        for (int i = 0; i < 30; i++) lengths[i] = 5;
        dist.build(lengths, 30);
      } else {                  // dynamic code (RFC 1951 3.2.7)
        const int nlen = (int)b.bits(5) + 257, ndist = (int)b.bits(5) + 1, ncode = (int)b.bits(4) + 4;
        if (b.err || nlen > 286 || ndist > 30) return KE_DEFLATE;
        for (int i = 0; i < 19; i++) lengths[i] = 0;
        for (int i = 0; i < ncode; i++) lengths[clOrder[i]] = (u8)b.bits(3);
        if (b.err) return KE_DEFLATE;
        HuffCode<30>& cl = dist;   // the code-length code (19 symbols) borrows the distance table
        if (!cl.build(lengths, 19)) return KE_DEFLATE;
        int idx = 0;
        while (idx < nlen + ndist) {
          const int sym = cl.decode(b);
          if (sym < 0) return KE_DEFLATE;
          if (sym < 16) lengths[idx++] = (u8)sym;
          else {
            int rep; u8 val = 0;
            if (sym == 16) { if (idx == 0) return KE_DEFLATE; val = lengths[idx - 1]; rep = 3 + (int)b.bits(2); }
            else if (sym == 17) rep = 3 + (int)b.bits(3);
            else rep = 11 + (int)b.bits(7);
            if (b.err || idx + rep > nlen + ndist) return KE_DEFLATE;
            while (rep--) lengths[idx++] = val;
          }
        }
        if (lengths[256] == 0) return KE_DEFLATE;   // no end-of-block code
        u8 dl[30]; for (int i = 0; i < ndist; i++) dl[i] = lengths[nlen + i];
        if (!lit.build(lengths, nlen)) return KE_DEFLATE;
        if (!dist.build(dl, ndist)) return KE_DEFLATE;
      }
      while (true) {
        int sym = lit.decode(b);
        if (sym < 0) return KE_DEFLATE;
        if (sym < 256) { if (dst) { if (out >= dstCap) return KE_DEFLATE; dst[out] = (u8)sym; } out++; }
        else if (sym == 256) break;
        else {
          sym -= 257; if (sym >= 29) return KE_DEFLATE;
          const u32 len = lenBase[sym] + b.bits(lenExtra[sym]);
          const int ds = dist.decode(b);
          if (ds < 0 || ds >= 30) return KE_DEFLATE;
          const u32 d = distBase[ds] + b.bits(distExtra[ds]);
          if (b.err || d > out) return KE_DEFLATE;
          if (dst) { if (out + len > dstCap) return KE_DEFLATE; for (u32 i = 0; i < len; i++) dst[out + i] = dst[out + i - d]; }
          out += len;
        }
        if (out > 0x7fffffffu) return KE_TOO_LARGE;
      }
    } else return KE_DEFLATE;
    if (last) break;
  }
  *produced = out; return KE_NONE;
}

HD u32 uleb_len(u64 v) { u32 n = 1; while (v >>= 7) n++; return n; }

// list[k] = batch index of the k-th DEFLATEd change (ascending). pass 0: outLen[k] = size of the inflated change
// (8 bytes magic + checksum, chunk type 1, LEB128 length, body) and the original range is kept (getChanges hands the
// original bytes back). pass 1: writes it at arena[extraStart + outOff[k]).
struct InflateKernel {
  int pass; u8* arena; u32* chOff; u32* chLen; const u32* list; u32* outLen; const u32* outOff; u32 extraStart; u32* origOff; u32* origLen; u64* errWord;
  HD void operator()(size_t k) const {
    const u32 c = list[k]; const u32 off = pass == 0 ? chOff[c] : origOff[k], len = pass == 0 ? chLen[c] : origLen[k];
    ByteReader r(arena, off + 9, off + len); const u64 clen = r.uleb();
    if (r.err || (u64)r.pos + clen > (u64)off + len) { raise(errWord, r.err ? r.err : (u32)KE_SUBARRAY, c); if (pass == 0) outLen[k] = 0; return; }
    if (pass == 0) {
      u32 n = 0; const u32 e = inflate_raw(arena, r.pos, r.pos + (u32)clen, nullptr, 0, &n);
      if (e) { raise(errWord, e, c); outLen[k] = 0; return; }
      outLen[k] = 9 + uleb_len(n) + n; origOff[k] = off; origLen[k] = len;
      return;
    }
    const u32 total = outLen[k]; if (total == 0) return;
    u8* dst = arena + extraStart + outOff[k];
    for (int i = 0; i < 8; i++) dst[i] = arena[off + i];
    dst[8] = 1;
    u32 hl = 9, n = 0;   // total = 9 + uleb_len(n) + n
    for (u32 w = 1; w <= 5; w++) { n = total - 9 - w; if (uleb_len(n) == w) break; }
    { u64 v = n; do { u8 x = v & 0x7f; v >>= 7; if (v) x |= 0x80; dst[hl++] = x; } while (v); }
    u32 produced = 0; const u32 e = inflate_raw(arena, r.pos, r.pos + (u32)clen, dst + hl, n, &produced);
    if (e || produced != n) { raise(errWord, e ? e : (u32)KE_DEFLATE, c); return; }
  }
};
// re-points the inflated changes (separate from pass 1: the batch-wide SHA kernel may still be reading the old entries)
struct InflatePatchKernel {
  const u32* list; const u32* outLen; const u32* outOff; u32 extraStart; u32* chOff; u32* chLen;
  HD void operator()(size_t k) const { if (outLen[k] == 0) return; const u32 c = list[k]; chOff[c] = extraStart + outOff[k]; chLen[c] = outLen[k]; }
};
struct DeflateFlagKernel {   // 1 for chunks of type 2 (columnar.js:742)
  const u8* arena; const u32* chOff; const u32* chLen; u32* flag;
  HD void operator()(size_t c) const { const u8* p = arena + chOff[c]; flag[c] = (chLen[c] > 8 && p[8] == 2 && p[0] == 0x85) ? 1u : 0u; }
};

}  // namespace amg
