// amgpu — kernels #5: column encoders for Backend.save().
//
// Replaces (reference paths relative to /root/reference):
//   backend/encoding.js:558-783   RLEEncoder (uint / int / utf8)     backend/encoding.js:932-998   DeltaEncoder
//   backend/encoding.js:1061-1135 BooleanEncoder                      backend/new.js:1680-1692      appendChange
//   backend/columnar.js:983-1004  encodeDocumentHeader (host side: engine_impl.cuh saveDocument)
//
// The reference appends value by value to stateful encoders; the result is the canonical run-length form of the whole
// value sequence (the decoder rejects every other form, encoding.js:860-913). That form is computed here for all
// values at once: equal neighbours -> runs (scan), runs -> records (null run | repetition | literal group of single
// values, scan), records -> byte sizes (scan) -> bytes. One templated kernel set serves every column type through a
// small value accessor (null test, equality, encoded size, write).
#pragma once
#include "opset.cuh"
#include "prims.cuh"
#include "gate.cuh"

namespace amg {

static const long long NULLV = (long long)0x8000000000000000ULL;
HD u32 uleb_size(u64 v) { u32 n = 1; while (v >>= 7) n++; return n; }
HD u32 sleb_size(long long v) { u32 n = 1; while (true) { const u32 b = (u32)(v & 0x7f); v >>= 7; if ((v == 0 && !(b & 0x40)) || (v == -1 && (b & 0x40))) return n; n++; } }
HD u32 put_uleb(u8* p, u64 v) { u32 n = 0; do { u8 b = v & 0x7f; v >>= 7; if (v) b |= 0x80; p[n++] = b; } while (v); return n; }
HD u32 put_sleb(u8* p, long long v) { u32 n = 0; while (true) { u8 b = (u8)(v & 0x7f); v >>= 7; if ((v == 0 && !(b & 0x40)) || (v == -1 && (b & 0x40))) { p[n++] = b; return n; } p[n++] = b | 0x80; } }

// ---- value accessors
struct NumCol {   // int64 values, NULLV = null; isSigned picks sLEB / uLEB raw values
  const long long* v; int isSigned;
  HD bool isNull(size_t i) const { return v[i] == NULLV; }
  HD bool eq(size_t i, size_t j) const { return v[i] == v[j]; }
  HD u32 size(size_t i) const { return isSigned ? sleb_size(v[i]) : uleb_size((u64)v[i]); }
  HD u32 write(size_t i, u8* dst) const { return isSigned ? put_sleb(dst, v[i]) : put_uleb(dst, (u64)v[i]); }
};
struct StrCol {   // byte strings in the arena, len == NULL32 = null
  const u8* arena; const u32* off; const u32* len;
  HD bool isNull(size_t i) const { return len[i] == NULL32; }
  HD bool eq(size_t i, size_t j) const {
    if (len[i] != len[j]) return false;
    if (len[i] == NULL32) return true;
    for (u32 k = 0; k < len[i]; k++) if (arena[off[i] + k] != arena[off[j] + k]) return false;
    return true;
  }
  HD u32 size(size_t i) const { return uleb_size(len[i]) + len[i]; }
  HD u32 write(size_t i, u8* dst) const { u32 n = put_uleb(dst, len[i]); for (u32 k = 0; k < len[i]; k++) dst[n + k] = arena[off[i] + k]; return n + len[i]; }
};

// ---- run-length pipeline
template <class V> struct EncHeadKernel { V c; u32* head; HD void operator()(size_t i) const { head[i] = (i == 0 || !c.eq(i, i - 1)) ? 1u : 0u; } };
struct EncRunStartKernel { const u32* head; const u32* headScan; u32* runStart; HD void operator()(size_t i) const { if (head[i]) runStart[headScan[i]] = (u32)i; } };
enum { RK_NULL = 0, RK_REP = 1, RK_SINGLE = 2 };
template <class V> struct EncRunKindKernel {   // record heads: every null run, every repetition, the first run of a literal group
  V c; const u32* runStart; u32* recHead;
  HD u32 kind(size_t r) const { if (c.isNull(runStart[r])) return RK_NULL; return runStart[r + 1] - runStart[r] >= 2 ? RK_REP : RK_SINGLE; }
  HD void operator()(size_t r) const { recHead[r] = (kind(r) != RK_SINGLE || r == 0 || kind(r - 1) != RK_SINGLE) ? 1u : 0u; }
};
struct EncRecFirstKernel { const u32* recHead; const u32* recScan; u32* recFirst; HD void operator()(size_t r) const { if (recHead[r]) recFirst[recScan[r]] = (u32)r; } };
template <class V> struct EncSizeKernel {
  V c; const u32* runStart; const u32* recHead; const u32* recScan; const u32* recFirst; u32* bytes;
  HD void operator()(size_t r) const {
    const u32 i = runStart[r], len = runStart[r + 1] - i; u32 b;
    if (c.isNull(i)) b = 1 + uleb_size(len);
    else if (len >= 2) b = sleb_size((long long)len) + c.size(i);
    else { b = c.size(i); if (recHead[r]) { const u32 rec = recScan[r]; b += sleb_size(-(long long)(recFirst[rec + 1] - recFirst[rec])); } }
    bytes[r] = b;
  }
};
template <class V> struct EncWriteKernel {
  V c; const u32* runStart; const u32* recHead; const u32* recScan; const u32* recFirst; const u32* byteOff; u8* out;
  HD void operator()(size_t r) const {
    const u32 i = runStart[r], len = runStart[r + 1] - i; u8* p = out + byteOff[r];
    if (c.isNull(i)) { p[0] = 0; put_uleb(p + 1, len); }
    else if (len >= 2) { const u32 n = put_sleb(p, (long long)len); c.write(i, p + n); }
    else { u32 n = 0; if (recHead[r]) { const u32 rec = recScan[r]; n = put_sleb(p, -(long long)(recFirst[rec + 1] - recFirst[rec])); } c.write(i, p + n); }
  }
};
// booleans: alternating run lengths starting with `false` (encoding.js:1061-1135)
struct BoolHeadKernel { const u32* v; u32* head; HD void operator()(size_t i) const { head[i] = (i == 0 || (v[i] != 0) != (v[i - 1] != 0)) ? 1u : 0u; } };
struct BoolSizeKernel { const u32* v; const u32* runStart; u32* bytes; HD void operator()(size_t r) const { bytes[r] = uleb_size(runStart[r + 1] - runStart[r]) + ((r == 0 && v[0] != 0) ? 1u : 0u); } };
struct BoolWriteKernel { const u32* v; const u32* runStart; const u32* byteOff; u8* out; HD void operator()(size_t r) const { u8* p = out + byteOff[r]; if (r == 0 && v[0] != 0) *p++ = 0; put_uleb(p, runStart[r + 1] - runStart[r]); } };
// delta columns: difference to the previous non-null value (encoding.js:960-975)
struct NonNullFlagKernel { const long long* v; u32* flag; HD void operator()(size_t i) const { flag[i] = v[i] != NULLV ? 1u : 0u; } };
struct CompactValKernel { const long long* v; const u32* flag; const u32* slot; long long* out; HD void operator()(size_t i) const { if (flag[i]) out[slot[i]] = v[i]; } };
struct DeltaKernel { const long long* v; const u32* flag; const u32* slot; const long long* compact; long long* out; HD void operator()(size_t i) const { if (!flag[i]) { out[i] = NULLV; return; } const u32 k = slot[i]; out[i] = v[i] - (k ? compact[k - 1] : 0); } };
// raw bytes: concatenation of per-row byte ranges
struct RawCopyKernel { const u8* arena; const u32* off; const u32* len; const u32* dstOff; u8* out; HD void operator()(size_t i) const { for (u32 k = 0; k < len[i]; k++) out[dstOff[i] + k] = arena[off[i] + k]; } };

// ---- values of the document columns (columnar.js:60-82) and of the change-metadata columns (columnar.js:84-94)
enum { SC_OBJ_ACTOR, SC_OBJ_CTR, SC_KEY_ACTOR, SC_KEY_CTR, SC_ID_ACTOR, SC_ID_CTR, SC_ACTION, SC_VAL_LEN, SC_SUCC_NUM };
struct SaveOpValKernel {   // one int64 value per document row
  int which; DocRows d; const u32* succOff; long long* out;
  HD void operator()(size_t p) const {
    long long v = NULLV;
    switch (which) {
      case SC_OBJ_ACTOR: if (d.obj[p]) v = id_actor(d.obj[p]); break;
      case SC_OBJ_CTR: if (d.obj[p]) v = (long long)id_ctr(d.obj[p]); break;
      case SC_KEY_ACTOR: if (d.keyStrLen[p] == NULL32 && d.key[p]) v = id_actor(d.key[p]); break;
      case SC_KEY_CTR: if (d.keyStrLen[p] == NULL32) v = (long long)id_ctr(d.key[p]); break;
      case SC_ID_ACTOR: v = id_actor(d.id[p]); break;
      case SC_ID_CTR: v = (long long)id_ctr(d.id[p]); break;
      case SC_ACTION: v = flags_action(d.flags[p]); break;
      case SC_VAL_LEN: v = d.valLen[p]; break;
      case SC_SUCC_NUM: v = succOff[p + 1] - succOff[p]; break;
    }
    out[p] = v;
  }
};
struct SaveSuccValKernel { int ctr; const u64* succ; long long* out; HD void operator()(size_t s) const { out[s] = ctr ? (long long)id_ctr(succ[s]) : (long long)id_actor(succ[s]); } };
struct SaveInsertKernel { DocRows d; u32* out; HD void operator()(size_t p) const { out[p] = d.flags[p] & F_INSERT; } };
struct SaveValBytesKernel { DocRows d; u32* out; HD void operator()(size_t p) const { out[p] = d.valLen[p] >> 4; } };
enum { SM_ACTOR, SM_SEQ, SM_MAX_OP, SM_TIME, SM_DEPS_NUM, SM_EXTRA_LEN };
struct SaveChangeValKernel {   // one int64 value per applied change, from its parsed header
  int which; const u8* arena; const ChangeMeta* meta; const ActorSlot* slots; u64 mask; long long* out; u32* strOff; u32* strLen; u64* errWord;
  HD void operator()(size_t c) const {
    const ChangeMeta& m = meta[c]; long long v = 0;
    switch (which) {
      case SM_ACTOR: { const u32 s = actor_find(slots, mask, fnv1a64(arena + m.actorOff, m.actorLen)); if (s == EMPTY32) { raise(errWord, KE_UNKNOWN_ACTOR, c); v = 0; } else v = slots[s].actorNum; } break;
      case SM_SEQ: v = (long long)m.seq; break;
      case SM_MAX_OP: v = (long long)(m.startOp + m.nOps) - 1; break;
      case SM_TIME: v = m.time; break;
      case SM_DEPS_NUM: v = m.nDeps; break;
      case SM_EXTRA_LEN: v = ((long long)m.extraLen << 4) | 7; strOff[c] = m.extraOff; strLen[c] = m.extraLen; break;
    }
    out[c] = v;
  }
};
struct SaveMessageKernel { const ChangeMeta* meta; u32* strOff; u32* strLen; HD void operator()(size_t c) const { strOff[c] = meta[c].msgOff; strLen[c] = meta[c].msgLen; } };
struct SaveDepIndexKernel { const u32* depIdx; long long* out; HD void operator()(size_t k) const { out[k] = depIdx[k]; } };

// change metadata of a loaded document: its columns are decoded again (one thread per column, as Backend.load does for
// the op columns) so that changes applied after the load can be appended before re-encoding
enum { LC_UINT, LC_DELTA, LC_STRING, LC_EXTRA_LEN, LC_SUM };
struct LoadedColKernel {
  int kind; const u8* arena; u32 off, len, rawOff; u32 count; long long* out; u32* strOff; u32* strLen; u64* sum;
  HD void operator()(size_t) const {
    RleReader r(arena, off, off + len, kind == LC_STRING ? 2 : (kind == LC_DELTA ? 1 : 0));
    long long acc = 0; u64 total = 0; u32 raw = rawOff;
    for (u32 i = 0; kind == LC_SUM ? !r.done() : i < count; i++) {
      long long n = 0; u32 o = 0, l = 0; const bool nn = r.next(n, o, l);
      if (kind == LC_SUM) { if (nn) total += (u64)n; continue; }
      if (kind == LC_STRING) { strOff[i] = o; strLen[i] = nn ? l : NULL32; continue; }
      if (kind == LC_DELTA) { if (nn) { acc += n; out[i] = acc; } else out[i] = NULLV; continue; }
      out[i] = nn ? n : NULLV;
      if (kind == LC_EXTRA_LEN) { const u32 bytes = nn ? (u32)((u64)n >> 4) : 0; strOff[i] = raw; strLen[i] = bytes; raw += bytes; }
    }
    if (kind == LC_SUM) *sum = total;
  }
};

struct ColumnEncoder {
  Ctx& ctx; ScanTemp& st;
  DBuf<u32> head, headScan, runStart, recHead, recScan, recFirst, bytes, byteOff, nnFlag, nnSlot; DBuf<long long> compact, delta;
  DBuf<u8> out; size_t outLen = 0;   // all columns of one save() back to back
  ColumnEncoder(Ctx& c, ScanTemp& s) : ctx(c), st(s) {}
  u32 readU32(const u32* p) { u32 v = 0; d2h(ctx, &v, p, 4); sync(ctx); return v; }
  void reserveOut(size_t extra) { out.ensure(ctx, outLen + extra + 64, outLen); }
  template <class V> size_t rle(const V& c, size_t n) {   // returns the encoded length; bytes are appended to `out`
    if (n == 0) return 0;
    head.ensure(ctx, n + 1); headScan.ensure(ctx, n + 2);
    foreach(ctx, n, EncHeadKernel<V>{c, head.p});
    scan_exclusive(ctx, st, head.p, headScan.p, n);
    const size_t R = readU32(headScan.p + n);
    runStart.ensure(ctx, R + 2); recHead.ensure(ctx, R + 1); recScan.ensure(ctx, R + 2); bytes.ensure(ctx, R + 1); byteOff.ensure(ctx, R + 2);
    foreach(ctx, n, EncRunStartKernel{head.p, headScan.p, runStart.p});
    { const u32 nn = (u32)n; h2d(ctx, runStart.p + R, &nn, 4); }
    if (R == 1) { nnFlag.ensure(ctx, 2); foreach(ctx, 1, FirstNullKernel<V>{c, nnFlag.p}); if (readU32(nnFlag.p)) return 0; }   // only nulls: nothing is written (encoding.js:778-782)
    foreach(ctx, R, EncRunKindKernel<V>{c, runStart.p, recHead.p});
    scan_exclusive(ctx, st, recHead.p, recScan.p, R);
    const size_t numRec = readU32(recScan.p + R);
    recFirst.ensure(ctx, numRec + 2);
    foreach(ctx, R, EncRecFirstKernel{recHead.p, recScan.p, recFirst.p});
    { const u32 rr = (u32)R; h2d(ctx, recFirst.p + numRec, &rr, 4); }
    foreach(ctx, R, EncSizeKernel<V>{c, runStart.p, recHead.p, recScan.p, recFirst.p, bytes.p});
    scan_exclusive(ctx, st, bytes.p, byteOff.p, R);
    const size_t total = readU32(byteOff.p + R);
    reserveOut(total);
    foreach(ctx, R, EncWriteKernel<V>{c, runStart.p, recHead.p, recScan.p, recFirst.p, byteOff.p, out.p + outLen});
    outLen += total; return total;
  }
  template <class V> struct FirstNullKernel { V c; u32* flag; HD void operator()(size_t) const { flag[0] = c.isNull(0) ? 1u : 0u; } };
  size_t rleNum(const long long* v, size_t n, bool isSigned) { return rle(NumCol{v, isSigned ? 1 : 0}, n); }
  size_t deltaNum(const long long* v, size_t n) {
    if (n == 0) return 0;
    nnFlag.ensure(ctx, n + 1); nnSlot.ensure(ctx, n + 2); compact.ensure(ctx, n + 1); delta.ensure(ctx, n + 1);
    foreach(ctx, n, NonNullFlagKernel{v, nnFlag.p});
    scan_exclusive(ctx, st, nnFlag.p, nnSlot.p, n);
    foreach(ctx, n, CompactValKernel{v, nnFlag.p, nnSlot.p, compact.p});
    foreach(ctx, n, DeltaKernel{v, nnFlag.p, nnSlot.p, compact.p, delta.p});
    return rleNum(delta.p, n, true);
  }
  size_t boolean(const u32* v, size_t n) {
    if (n == 0) return 0;
    head.ensure(ctx, n + 1); headScan.ensure(ctx, n + 2);
    foreach(ctx, n, BoolHeadKernel{v, head.p});
    scan_exclusive(ctx, st, head.p, headScan.p, n);
    const size_t R = readU32(headScan.p + n);
    runStart.ensure(ctx, R + 2); bytes.ensure(ctx, R + 1); byteOff.ensure(ctx, R + 2);
    foreach(ctx, n, EncRunStartKernel{head.p, headScan.p, runStart.p});
    { const u32 nn = (u32)n; h2d(ctx, runStart.p + R, &nn, 4); }
    foreach(ctx, R, BoolSizeKernel{v, runStart.p, bytes.p});
    scan_exclusive(ctx, st, bytes.p, byteOff.p, R);
    const size_t total = readU32(byteOff.p + R);
    reserveOut(total);
    foreach(ctx, R, BoolWriteKernel{v, runStart.p, byteOff.p, out.p + outLen});
    outLen += total; return total;
  }
  size_t raw(const u8* arena, const u32* off, const u32* len, size_t n) {
    if (n == 0) return 0;
    byteOff.ensure(ctx, n + 2);
    scan_exclusive(ctx, st, len, byteOff.p, n);
    const size_t total = readU32(byteOff.p + n);
    reserveOut(total);
    foreach(ctx, n, RawCopyKernel{arena, off, len, byteOff.p, out.p + outLen});
    outLen += total; return total;
  }
};

}  // namespace amg
