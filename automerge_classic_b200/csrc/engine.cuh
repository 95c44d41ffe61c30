// amgpu — host orchestration of the change-replay pipeline (one Engine = one backend document).
//
// Mirrors class BackendDoc (reference backend/new.js:1694-2069): applyChanges (:1797-1879) and
// getPatch (:2060-2068) run as sequences of CUDA kernels over device-resident state:
//   arena   : every applied / queued change's bytes back to back (values and keys are referenced in place)
//   hashes  : SHA-256 of every applied change                      (gate)
//   doc     : document rows in document order, SoA, + succ CSR      (op set)
//   actors  : byte-string table of actor ids -> document actor index
// An applyChanges call is atomic (reference :1793-1795): all kernels write into scratch buffers; the
// persistent state is swapped in only after the last error check passed.
#pragma once
#include <algorithm>
#include <array>
#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <thread>
#include <zlib.h>
#include "patch.cuh"
#include "inflate.cuh"
#include "encode.cuh"
#include "prims.cuh"
#include "domlocal.cuh"
#include "doccols.cuh"
#include "history.cuh"
#include "unknowncols.hpp"

namespace amg {

struct DocBufs {
  DBuf<u64> id, obj, key; DBuf<u32> keyStrOff, keyStrLen, flags, valLen, valOff, time;
  void ensure(Ctx& c, size_t n, size_t keep = 0) {
    id.ensure(c, n, keep); obj.ensure(c, n, keep); key.ensure(c, n, keep); keyStrOff.ensure(c, n, keep); keyStrLen.ensure(c, n, keep);
    flags.ensure(c, n, keep); valLen.ensure(c, n, keep); valOff.ensure(c, n, keep); time.ensure(c, n, keep);
  }
  DocRows view() { return DocRows{id.p, obj.p, key.p, keyStrOff.p, keyStrLen.p, flags.p, valLen.p, valOff.p, time.p}; }
  void swap(DocBufs& o) {
    std::swap(id.p, o.id.p); std::swap(id.cap, o.id.cap); std::swap(obj.p, o.obj.p); std::swap(obj.cap, o.obj.cap); std::swap(key.p, o.key.p); std::swap(key.cap, o.key.cap);
    std::swap(keyStrOff.p, o.keyStrOff.p); std::swap(keyStrOff.cap, o.keyStrOff.cap); std::swap(keyStrLen.p, o.keyStrLen.p); std::swap(keyStrLen.cap, o.keyStrLen.cap);
    std::swap(flags.p, o.flags.p); std::swap(flags.cap, o.flags.cap); std::swap(valLen.p, o.valLen.p); std::swap(valLen.cap, o.valLen.cap);
    std::swap(valOff.p, o.valOff.p); std::swap(valOff.cap, o.valOff.cap); std::swap(time.p, o.time.p); std::swap(time.cap, o.time.cap);
  }
};

struct PatchOut {   // flat patch, written straight into the engine's pinned output buffer (layout: include/amgpu.h)
  u64 maxOp = 0, pendingChanges = 0; bool hasActorSeq = false; std::string actor; u64 seq = 0;
  std::vector<std::pair<u32, u64>> clock; std::vector<std::array<u8, 32>> deps; std::vector<std::string> actors;
  size_t propsOff = 0, numProps = 0, editsOff = 0, numEdits = 0, elemOff = 0, valBytesOff = 0, valBytesLen = 0, bigEnd = 0;   // byte offsets into Engine::patchBuf
  const u8* bytes = nullptr; size_t bytesLen = 0;   // final serialised patch (valid until the next call on the same engine)
};

struct HostChange { u32 off, len; };   // (arena offset, length) of the inflated change

// Pinned host mirror of the arena: grows without zero-filling; H2D copies read straight from it.
struct HostArena {
  HBuf<u8> buf; size_t len = 0;
  u8* data() { return buf.p; } const u8* data() const { return buf.p; }
  size_t size() const { return len; }
  void reserve(size_t n) { buf.ensure(n + 64); }
  void resize(size_t n) { if (n > len) buf.ensure(n + 64); len = n; }
  void append(const void* p, size_t n) { const size_t at = len; resize(len + n); memcpy(buf.p + at, p, n); }
  void assign(const HostArena& o) { resize(o.len); if (o.len) memcpy(buf.p, o.buf.p, o.len); }
};

inline std::string hex_of(const u8* p, size_t n) { static const char* d = "0123456789abcdef"; std::string s; for (size_t i = 0; i < n; i++) { s.push_back(d[p[i] >> 4]); s.push_back(d[p[i] & 15]); } return s; }

class Engine {
 public:
  Ctx ctx;
  // ---- persistent device state
  DBuf<u8> arena; size_t arenaLen = 0; HostArena hostArena;   // pinned host mirror (values / keys / changes are read from it)
  DBuf<u8> hashes; size_t numApplied = 0;
  DocBufs doc; size_t numRows = 0; DBuf<u32> succOff; DBuf<u64> succ; size_t numSucc = 0;
  DBuf<ActorSlot> actorSlots; size_t actorCap = 0;
  DBuf<u32> actorRank;
  // ---- persistent host state
  std::vector<std::string> actorIds;   // raw bytes, index = document actor number
  std::vector<u64> clock;              // per actor number
  std::vector<std::array<u8, 32>> heads; std::vector<u32> headIdx;   // heads (sorted by hash) and their application indices
  std::vector<std::pair<u32, u32>> actorRep;   // arena (offset, length) of each actor's id bytes
  std::vector<HostChange> changes;     // applied, in application order
  std::vector<std::array<u8, 32>> changeHashes;   // host copy of applied hashes (filled lazily)
  struct OrigRange { u32 idx; HostChange range; };
  std::vector<OrigRange> deflatedOriginal;        // (applied change index, arena range of the original DEFLATEd bytes), ascending index
  const HostChange* originalOf(u32 idx) const {
    auto it = std::lower_bound(deflatedOriginal.begin(), deflatedOriginal.end(), idx, [](const OrigRange& a, u32 v) { return a.idx < v; });
    return it != deflatedOriginal.end() && it->idx == idx ? &it->range : nullptr;
  }
  std::vector<HostChange> queue, queueOriginal;   // not yet causally ready (+ original range, len 0 = not deflated)
  u64 maxOp = 0;
  float lastPhaseMs[24] = {0};   // [0..11] CUDA-event phases, [12..23] host wall-clock markers (ms since call start)
  struct PhaseTimer* curTimer = nullptr; std::function<void()> curHostMark;
  std::vector<std::pair<const char*, float>> dbgMarks; std::function<void(const char*)> dbgMark = [](const char*) {};
  HBuf<u8> patchBuf;   // pinned: patch records are copied device -> host directly into their final place
  // ---- scratch (grow-only)
  DBuf<u32> chOff, chLen, nOps, nPreds, nDeps, nActors, colOff, colLen, depBase, depIdx, primary, pass, flagWord, appRank, opBase, predBase, timeBase, amapBase, amap, authorSlot, newSlots;
  DBuf<u8> applied; DBuf<ChangeHot> hot; DBuf<ChangeMeta> meta /* save(): full headers */; DBuf<u64> errWord; DBuf<u32> hashTable;
  DBuf<u32> inflCap, inflCapOff, inflOvf; DBuf<u8> inflScratch;
  DBuf<u32> rawBase, rawPredBase, decErr, decDirect; DBuf<u64> decCursor; size_t lastDeflCount = 0, lastDeflStart = 0;
  u32* decTotalsPtr() { return reinterpret_cast<u32*>(decCursor.p + 2); }
  DBuf<u32> patchByteLen, patchByteOff; DBuf<u8> patchBytesD;   // key / value bytes shipped inside the patch   // fused decode (decode.cuh k_decode_tiles)
  DBuf<u32> r_objActor, r_objCtr, r_keyActor, r_keyCtr, r_keyStrOff, r_keyStrLen, r_insert, r_action, r_valLen, r_valOff, r_predNum, r_predOff, r_predActor, r_predCtr;
  DBuf<u64> o_id, o_obj, o_key, o_predId; DBuf<u32> o_keyStrOff, o_keyStrLen, o_flags, o_valLen, o_valOff, o_predOff, o_predNum, o_change, o_time;
  DBuf<u32> isRow, rowSlot, rowOfOp; DocBufs work, sorted;
  DBuf<u64> idKeys; DBuf<u32> idVals; DBuf<u32> objRow, elemRow, parentRow, keySlot, repList, repCount, listPos, perm, pos;
  DBuf<KeySlot> keySlots; DBuf<u64> sortKeys; DBuf<u32> sortVals; SortTemp sortTmp; ScanTemp scanTmp;
  ParColumnDecoder parCols{ctx, scanTmp}; size_t parDocMinRows = 4096;   // documents with at least this many rows decode their columns in parallel (doccols.cuh); AMG_PAR_DOC_MIN overrides
  DBuf<u32> eNext, eRank, insItems, itemIdx, objSlot; DBuf<u64> ePacked, ePacked2;
  DBuf<u64> pairKey, pairSucc, newSucc; DBuf<u32> pairIdx, pairPos, pairTime, succCnt, newSuccCnt, newSuccOff, firstNewSucc;
  DBuf<u32> elemPos, keyRankAt, objPos, head, headScan, groupOf, groupRows, groupVisible, groupFirst, groupTouched, groupLinked, objTouchedAt, linkDone, emit, marker, slot;
  DBuf<u32> isObjHead, objIdx, objStart, elemVis, elemVisScan, rowEmit, firstVis, state, nItems, itemBase, qIndex, zero, wzero, zscan, wscan, editObjKey;
  DBuf<DomItem> items, items2; DBuf<PropRec> propOut; DBuf<EditRec> editOut, editOut2; DBuf<u64> editElem, editElem2;
  std::unique_ptr<ColumnEncoder> encoder; DBuf<long long> saveVals; DBuf<u32> saveStrOff, saveStrLen; std::string loadedDoc; size_t numLoaded = 0; HostChange loadedCols[9] = {}; DBuf<u64> counterTotal; DBuf<int> domW, domW2; DBuf<u32> elemMinT, editRowPos, editRowPos2, editObjKey2, rowClass, firstBare, counterOwner, newSuccTime, counterLast, runHeadFlag, runScan, runStart, elemFollower, domTw, domTw2, oldVisScan, inflLen, inflOff, groupHasChild, gCount, gElem, gT1, gQOrd, gBase, nQ, elemHasRecs, listLinkTime, editElemPos, editElemPos2, editKind, editPred, editDead, editMerge, editMulti, editLive;
  DBuf<u32> seqSlot, actorCnt, actorBaseD, clockD, changeActor, editTime; DBuf<u8> hashTmp; DBuf<u32> headsPack, headsOut; bool batchInOrder = true;
  DBuf<u32> finalTime, gFailed, memberFinal, opAt, runHead, opGroupHead; DBuf<u64> gBound; DocRows workView{}; DBuf<HostChange> chPairs; DBuf<u32> largeFlag, largeSlot, largeList; size_t lastNumLarge = 0; DBuf<u64> zwScan; DBuf<u32> deflList, patchTriples;

  explicit Engine(int device) {
    ctx.device = device;
#ifndef AMG_EMU
    int count = 0; cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) throw Error(AMG_ERR_CUDA, "amgpu: no CUDA device available (this library has no CPU fallback)");
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop; CUDA_CHECK(cudaGetDeviceProperties(&prop, device)); ctx.numSMs = prop.multiProcessorCount;
    CUDA_CHECK(cudaStreamCreateWithFlags(&ctx.stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&ctx.side, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&ctx.copy, cudaStreamNonBlocking));
    ctx.peekCap = 1u << 20; CUDA_CHECK(cudaMallocHost((void**)&ctx.peekBuf, ctx.peekCap + 64));
    ctx.peekFlag = reinterpret_cast<volatile unsigned long long*>(ctx.peekBuf + ctx.peekCap); *ctx.peekFlag = 0;
    CUDA_CHECK(cudaEventCreateWithFlags(&ctx.evUp, cudaEventDisableTiming)); CUDA_CHECK(cudaEventCreateWithFlags(&ctx.evMirror, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&ctx.evFork, cudaEventDisableTiming)); CUDA_CHECK(cudaEventCreateWithFlags(&ctx.evJoin, cudaEventDisableTiming));
    ShaConsts k; memcpy(k.k, SHA_K, sizeof(SHA_K)); CUDA_CHECK(cudaMemcpyToSymbol(c_sha, &k, sizeof(k)));
#endif
    errWord.ensure(ctx, 4); flagWord.ensure(ctx, 16);
    actorCap = 64; actorSlots.ensure(ctx, actorCap); resetActorSlots(0, actorCap);
    succOff.ensure(ctx, 1); dev_memset(ctx, succOff.p, 0, 4);
  }
  ~Engine() {
#ifndef AMG_EMU
    if (ctx.stream) cudaStreamDestroy(ctx.stream);
    mirror_wait(ctx);
    if (ctx.side) cudaStreamDestroy(ctx.side);
    if (ctx.copy) cudaStreamDestroy(ctx.copy);
    if (ctx.peekBuf) cudaFreeHost(ctx.peekBuf);
    if (ctx.phaseEvReady) for (auto& e : ctx.phaseEv) cudaEventDestroy(e);
    if (last_peek_ctx() == &ctx) last_peek_ctx() = nullptr;
    if (ctx.evUp) cudaEventDestroy(ctx.evUp);
    if (ctx.evMirror) cudaEventDestroy(ctx.evMirror);
    if (ctx.evFork) cudaEventDestroy(ctx.evFork);
    if (ctx.evJoin) cudaEventDestroy(ctx.evJoin);
#endif
  }

  // (re)builds the device actor table from the host's actor list (after growth, rollback or commit)
  void rebuildActorTable() {
    std::vector<ActorSlot> t(actorCap); for (auto& s : t) { s.hash = 0; s.first = ~0ULL; s.actorNum = EMPTY32; s.repOff = 0; s.repLen = 0; s.pad = 0; }
    for (size_t a = 0; a < actorIds.size(); a++) {
      const u64 h = fnv1a64((const u8*)actorIds[a].data(), (u32)actorIds[a].size()); u64 s = mix64(h) & (actorCap - 1);
      while (t[s].hash != 0) s = (s + 1) & (actorCap - 1);
      t[s].hash = h; t[s].first = 0; t[s].actorNum = (u32)a; t[s].repOff = actorRep[a].first; t[s].repLen = actorRep[a].second;
    }
    h2d(ctx, actorSlots.p, t.data(), actorCap * sizeof(ActorSlot)); sync(ctx);
  }
  void resetActorSlots(size_t from, size_t to) {
    std::vector<ActorSlot> init(to - from); for (auto& s : init) { s.hash = 0; s.first = ~0ULL; s.actorNum = EMPTY32; s.repOff = 0; s.repLen = 0; s.pad = 0; }
    h2d(ctx, actorSlots.p + from, init.data(), init.size() * sizeof(ActorSlot)); sync(ctx);
  }

  // ---------------------------------------------------------------- error plumbing
  // Every small device -> host read also brings the error word along (same sync); a later error check is free when no
  // kernel was launched in between.
  u64 errSnapshot = 0; uint64_t errSnapLaunches = ~0ull;
  u64* pinnedWords(size_t n) { hostWord.ensure(n + 1); return hostWord.p; }
  void readWords(std::initializer_list<std::pair<const void*, size_t>> srcs, void* const* dsts) {   // one sync for all of them + the error word
    u64* w = pinnedWords(srcs.size() + 1); size_t k = 0;
    const void* from[9]; size_t sizes[9]; void* to[9];
    if (srcs.size() > 8) throw Error(AMG_ERR_INTERNAL, "readWords: too many words");
    for (auto& s : srcs) { w[k] = 0; from[k] = s.first; sizes[k] = s.second; to[k] = &w[k]; k++; }
    from[k] = errWord.p; sizes[k] = 8; to[k] = &w[k];
#ifndef AMG_EMU
    ctx.peekFlagArmed = false;
#endif
    if (k + 1 <= 8) d2h_words(ctx, (int)k + 1, from, sizes, to); else for (size_t i = 0; i <= k; i++) d2h(ctx, to[i], from[i], sizes[i]);
    const uint64_t launchesNow = ctx.launches;
    sync(ctx, true);   // the words' kernel is the last thing on the stream: wait on its completion flag
    k = 0; for (auto& s : srcs) { memcpy(dsts[k], &w[k], s.second); k++; }
    errSnapshot = w[k]; errSnapLaunches = launchesNow;
  }
  u64 fetchErr() { if (errSnapLaunches != ctx.launches) { void* none[1] = {nullptr}; readWords({}, none); } return errSnapshot; }
  std::string opIdText(u64 id) const {
    const u32 a = id_actor(id);
    return std::to_string(id_ctr(id)) + "@" + (a < actorIds.size() ? hex_of((const u8*)actorIds[a].data(), actorIds[a].size()) : std::string("?"));
  }
  [[noreturn]] void throwKernelError(u64 w, const std::vector<std::string>& actorsNow, const u64* predIdHost = nullptr) {
    const u32 code = (u32)(w & 0xff); const u64 item = w >> 8; (void)item; (void)actorsNow; (void)predIdHost;
    switch (code) {
      case KE_MAGIC: throw Error(AMG_ERR_RANGE, "Data does not begin with magic bytes 85 6f 4a 83");
      case KE_CHECKSUM: throw Error(AMG_ERR_RANGE, "checksum does not match data");
      case KE_TRAILING: throw Error(AMG_ERR_RANGE, "Encoded change has trailing data");
      case KE_CHUNK_TYPE: throw Error(AMG_ERR_RANGE, "Unexpected chunk type");
      case KE_DEFLATE: throw Error(AMG_ERR_RANGE, "invalid deflate data");
      case KE_HIST_RANGE: throw Error(AMG_ERR_RANGE, "Operation ID outside of allowed range");        // columnar.js:925
      case KE_HIST_OPID: throw Error(AMG_ERR_RANGE, "Expected opId does not match the operation found");   // columnar.js:936
      case KE_HIST_DEP: throw Error(AMG_ERR_RANGE, "No hash for dependency index");                      // columnar.js:952
      case KE_TRUNCATED: throw Error(AMG_ERR_RANGE, "buffer ended with incomplete number");
      case KE_SUBARRAY: throw Error(AMG_ERR_RANGE, "subarray exceeds buffer size");
      case KE_NUM_RANGE: throw Error(AMG_ERR_RANGE, "number out of range");
      case KE_COL_ORDER: throw Error(AMG_ERR_RANGE, "Columns must be in ascending order");
      case KE_COL_DEFLATE: throw Error(AMG_ERR_RANGE, "change must not contain deflated columns");
      case KE_RLE_REP1: throw Error(AMG_ERR_RANGE, "Repetition count of 1 is not allowed, use a literal instead");
      case KE_RLE_SUCC_REP: throw Error(AMG_ERR_RANGE, "Successive repetitions with the same value are not allowed");
      case KE_RLE_SUCC_LIT: throw Error(AMG_ERR_RANGE, "Successive literals are not allowed");
      case KE_RLE_SUCC_NULL: throw Error(AMG_ERR_RANGE, "Successive null runs are not allowed");
      case KE_RLE_ZERO_NULL: throw Error(AMG_ERR_RANGE, "Zero-length null runs are not allowed");
      case KE_RLE_LIT_REP: throw Error(AMG_ERR_RANGE, "Repetition of values is not allowed in literal");
      case KE_BOOL_ZERO: throw Error(AMG_ERR_RANGE, "Zero-length runs are not allowed");
      case KE_OBJ_MISMATCH: throw Error(AMG_ERR_RANGE, "Mismatched object reference");
      case KE_KEY_MISMATCH: throw Error(AMG_ERR_RANGE, "Mismatched operation key");
      case KE_ACTOR_INDEX: throw Error(AMG_ERR_RANGE, "actor index out of range");
      case KE_TOO_LARGE: throw Error(AMG_ERR_UNSUPPORTED, "amgpu: value exceeds the engine's 32-bit counter / 4 GiB arena limits");
      case KE_UNKNOWN_ACTOR: throw Error(AMG_ERR_RANGE, "actorId is not known to document");
      case KE_PRED_MISSING: throw Error(AMG_ERR_RANGE, "no matching operation for pred");
      case KE_REF_ELEM: throw Error(AMG_ERR_RANGE, "Reference element not found");
      case KE_LIST_ELEM: throw Error(AMG_ERR_RANGE, "could not find list element with ID");
      case KE_DUP_OPID: throw Error(AMG_ERR_RANGE, "duplicate operation ID");
      case KE_LAMPORT: throw Error(AMG_ERR_UNSUPPORTED, "amgpu: insert with an opId not greater than its reference element (Lamport order violated)");
      case KE_HASH_COLLISION: throw Error(AMG_ERR_UNSUPPORTED, "amgpu: 64-bit string hash collision");
      case KE_GROUP_COLUMN: throw Error(AMG_ERR_UNSUPPORTED, "amgpu: a GROUP_CARD column of an unknown version in the group of known columns (the reference would read those as arrays)");
      case KE_UNSUPPORTED_OP: throw Error(AMG_ERR_UNSUPPORTED, "amgpu: operation pattern outside the incremental-patch subset (see DESIGN.md)");
      default: throw Error(AMG_ERR_INTERNAL, "amgpu: kernel error " + std::to_string(code));
    }
  }
  void checkErr(const std::vector<std::string>& actorsNow) { u64 w = fetchErr(); if (w) throwKernelError(w, actorsNow); }

  // ---------------------------------------------------------------- helpers
  DBuf<u64> gateBest;   // causal gate with several copies of a waiting change: best (pass, position) per hash
  bool domLocalReady = false;   // k_dom_local's dynamic shared memory size has been raised on this device
  std::vector<HostChange> batchStore;   // applyChanges: (offset, length) of the batch entries
  HBuf<HostChange> pairStage;           // the same table in pinned memory: uploaded by DMA
  HBuf<u32> pinnedScratch; HBuf<u64> hostWord;   // pinned landing slots for the small device -> host reads that size the next stage
  u32 readU32(const u32* dptr) { u32 v = 0; void* d[1] = {&v}; readWords({{dptr, 4}}, d); return v; }
  void readU32x2(const u32* a, const u32* b, u32* va, u32* vb) { void* d[2] = {va, vb}; readWords({{a, 4}, {b, 4}}, d); }
  void fill32(u32* p, u32 v, size_t n) { foreach(ctx, n, FillU32Kernel{p, v}); }
  // sort `perm` (row ids) by successive 64-bit fields produced by keyFn(field) ; stable LSD over fields
  void sortPairs(DBuf<u64>& keys, DBuf<u32>& vals, size_t n, int bits) { radix_sort_pairs(ctx, sortTmp, keys, vals, n, 0, bits); }

  // reference columnar.js:813-823 (pako.inflateRaw -> zlib raw inflate); magic + checksum are kept. `zs` may be a reusable,
  // already initialised stream (inflateInit2(.., -15)); it is reset, not re-allocated.
  static std::string inflateChange(const u8* buf, size_t len, z_stream* reuse = nullptr) {
    ByteReader r(buf, 9, (u32)len); const u64 clen = r.uleb();
    if (r.err || r.pos + clen > len) throw Error(AMG_ERR_RANGE, "buffer ended with incomplete number");
    z_stream local; z_stream* zs = reuse;
    if (!zs) { memset(&local, 0, sizeof(local)); if (inflateInit2(&local, -15) != Z_OK) throw Error(AMG_ERR_INTERNAL, "inflateInit failed"); zs = &local; }
    else inflateReset(zs);
    struct End { z_stream* z; bool own; ~End() { if (own) inflateEnd(z); } } end{zs, !reuse};
    std::string out; out.resize(std::max<size_t>(clen * 4, 512));
    zs->next_in = (Bytef*)(buf + r.pos); zs->avail_in = (uInt)clen; size_t produced = 0;
    while (true) {
      zs->next_out = (Bytef*)out.data() + produced; zs->avail_out = (uInt)(out.size() - produced);
      int rc = inflate(zs, Z_NO_FLUSH); produced = out.size() - zs->avail_out;
      if (rc == Z_STREAM_END) break;
      if (rc != Z_OK && rc != Z_BUF_ERROR) throw Error(AMG_ERR_RANGE, "invalid deflate data");
      if (zs->avail_out == 0) out.resize(out.size() * 2); else if (zs->avail_in == 0) throw Error(AMG_ERR_RANGE, "unexpected end of deflate data");
    }
    // header: magic + checksum (8 bytes), chunk type 1, LEB128 length, then the inflated body
    u8 hdr[24]; memcpy(hdr, buf, 8); hdr[8] = 1; size_t hl = 9; u64 v = produced; do { u8 b = v & 0x7f; v >>= 7; if (v) b |= 0x80; hdr[hl++] = b; } while (v);
    std::string res; res.reserve(hl + produced); res.append((const char*)hdr, hl); res.append(out.data(), produced);
    return res;
  }

  // ---------------------------------------------------------------- applyChanges
  struct ApplyResult { PatchOut patch; };

  void applyChanges(const u8* const* bufs, const size_t* lens, size_t n, const u8* blob, const u64* offsets, bool isLocal, bool wantPatch, PatchOut& out, bool hostScan = false);
  void applyChangesOnce(const u8* const* bufs, const size_t* lens, size_t n, const u8* blob, const u64* offsets, bool isLocal, bool wantPatch, PatchOut& out, bool hostScan = false);
  void getPatch(PatchOut& out);
  void saveDocument(std::string& result);
  void buildPatch(DocRows d, size_t N, bool wholeDoc, const OpRows* ops, size_t numOps, const IdTable* idt, const u32* rowOfOpD, const u32* posD,
                  const std::vector<std::string>& actorsNow, PatchOut& out, const u32* succOffD, const u64* succD);
  void fillPatchHeader(PatchOut& out);
  void finishPatch(PatchOut& out);
  void reset();
  void loadDocument(const u8* buf, size_t len);
  size_t historyRebuilt = 0;   // changes [0, historyRebuilt) were rebuilt by computeHashGraph (getChanges DEFLATEs the large ones like encodeChange does)
  DBuf<u64> excl64, offsDev;
  bool headIndexesUnknown = false;   // Backend.load of a document with several heads and no head indexes, until computeHashGraph has matched them
  bool haveHashGraph = true;   // false after Backend.load: change history (hashes, bytes) is not reconstructed (new.js:1887-1912)
  void benchDecode(int iters, float* msSha, float* msParse, float* msDec, u64* algoBytes);
  UnknownStore unknownCols;   // values of columns with ids this version does not know, per op (unknowncols.hpp)
  void collectUnknownColumns(size_t B, std::vector<std::pair<u64, UnknownRow>>& out, std::set<u32>& ids);
  void appendUnknownDocColumns(std::vector<std::pair<u32, std::string>>& cols);   // save(): their document columns
  RawRows rawRows();
  u32 decodeHugeChanges(const RawRows& raw, size_t numLarge); DBuf<u32> hugeDone;
  void runDecodeTiles(const u8* arenaP, size_t B, size_t batchBytes, const u32* deflListP = nullptr, size_t numDefl = 0, size_t deflStart = 0);
  DecodeTilesArgs decodeArgs(const u8* arenaP, size_t B, size_t batchBytes);
  // Host mirror of the arena, filled on demand: hostArena holds arena[0, hostArena.size()); whatever is missing is fetched
  // from the device when a reader (getChanges, amg_arena, clone ...) asks for it.
  void ensureHostMirror() {
    if (hostArena.size() >= arenaLen) return;
    const size_t from = hostArena.size(); hostArena.resize(arenaLen);
    d2h(ctx, hostArena.data() + from, arena.p + from, arenaLen - from); sync(ctx);
  }   // sizes the raw row tables and launches the fused decode
  bool decodeOverflowed(const u32 totals[4]);
  void computeHashGraph();   // change history of a loaded document (history.cuh)
  int debugDecodeColumn(const u8* bytes, size_t len, int kind, size_t n, bool parallel, long long* out);
  void decodeRaw(const u8* blob, const u64* offsets, size_t n, u8* hashesOut, u32* nOpsOut, u32** rowsOut, size_t* totalOps, size_t* totalPreds);
  size_t lastB = 0, lastM = 0, lastP = 0, lastBytes = 0;
  size_t decWantRows = 0, decWantPreds = 0, decRowCap = 0, decPredCap = 0;
};

}  // namespace amg
