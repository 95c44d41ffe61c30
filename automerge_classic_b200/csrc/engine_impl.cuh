// amgpu — Engine::applyChanges / getPatch pipeline (see engine.cuh for the state layout).
#pragma once
#include <chrono>
#ifndef AMG_EMU
#include <nvtx3/nvToolsExt.h>
#endif
#include "engine.cuh"
#include "misc.cuh"

namespace amg {

#ifdef AMG_EMU
#define CUDA_CHECK_EMU(x) do {} while (0)
#else
#define CUDA_CHECK_EMU(x) CUDA_CHECK(x)
#endif
static const size_t PATCH_HDR_WORDS = 20;
// NVTX range per pipeline phase: next() closes the running range and opens the named one (nullptr: just closes)
struct NvtxPhases {
#ifndef AMG_EMU
  bool open = false;
  void next(const char* name) { if (open) nvtxRangePop(); open = name != nullptr; if (name) nvtxRangePushA(name); }
  ~NvtxPhases() { if (open) nvtxRangePop(); }
#else
  void next(const char*) {}
#endif
};
struct HostClock {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  float ms() const { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};
struct PhaseTimer {
#ifndef AMG_EMU
  cudaEvent_t* ev; int n = 0; Ctx* c;   // the events live in the context: creating and destroying 13 timing events per call cost more than a pipeline phase
  explicit PhaseTimer(Ctx& ctx) : ev(ctx.phaseEv), c(&ctx) { if (!ctx.phaseEvReady) { for (int i = 0; i < 13; i++) cudaEventCreate(&ev[i]); ctx.phaseEvReady = true; } mark(); }
  void mark() { if (n < 13) cudaEventRecord(ev[n++], c->stream); }
  void collect(float* out, int maxN) { cudaEventSynchronize(ev[n - 1]); for (int i = 0; i + 1 < n && i < maxN; i++) cudaEventElapsedTime(&out[i], ev[i], ev[i + 1]); }
#else
  explicit PhaseTimer(Ctx&) {}
  void mark() {}
  void collect(float*, int) {}
#endif
};

inline void parallel_copy(u8* dst, const u8* src, size_t n) {
  if (n < (4u << 20)) { memcpy(dst, src, n); return; }
  unsigned nt = std::min<unsigned>(8, std::max(1u, std::thread::hardware_concurrency()));
  std::vector<std::thread> ts; size_t per = (n + nt - 1) / nt;
  for (unsigned t = 0; t < nt; t++) { size_t a = t * per, b = std::min(n, a + per); if (a < b) ts.emplace_back([=] { memcpy(dst + a, src + a, b - a); }); }
  for (auto& t : ts) t.join();
}

// changes [c0, c1) of a pointer array, back to back into dst (the shape Backend.applyChanges(state, Uint8Array[]) hands over)
// (a thread's loop asks for the buffers a few changes ahead - the sources are scattered heap objects, every one a cache miss)
inline void gather_range(u8* q, const u8* const* bufs, const size_t* lens, size_t a, size_t b) {
  for (size_t i = a; i < b; i++) {
    if (i + 8 < b) { __builtin_prefetch(bufs[i + 8]); __builtin_prefetch(bufs[i + 8] + 64); }
    memcpy(q, bufs[i], lens[i]); q += lens[i];
  }
}
inline void parallel_gather(u8* dst, const u8* const* bufs, const size_t* lens, size_t c0, size_t c1) {
  size_t bytes = 0; for (size_t i = c0; i < c1; i++) bytes += lens[i];
  unsigned nt = bytes < (4u << 20) ? 1u : std::min<unsigned>(16, std::max(1u, std::thread::hardware_concurrency() / 2));   // small scattered buffers: bound by cache misses, not by bandwidth
  if (nt == 1) { gather_range(dst, bufs, lens, c0, c1); return; }
  std::vector<std::thread> ts; const size_t per = (c1 - c0 + nt - 1) / nt; size_t at = 0;
  for (unsigned t = 0; t < nt; t++) {
    const size_t a = c0 + t * per, b = std::min(c1, a + per); if (a >= b) break;
    u8* d = dst + at; for (size_t i = a; i < b; i++) at += lens[i];
    ts.emplace_back([=] { gather_range(d, bufs, lens, a, b); });
  }
  for (auto& t : ts) t.join();
}

inline void Engine::fillPatchHeader(PatchOut& out) {
  out.maxOp = maxOp; out.pendingChanges = queue.size();
  out.clock.clear(); for (size_t a = 0; a < clock.size(); a++) if (clock[a] > 0) out.clock.emplace_back((u32)a, clock[a]);
  out.deps = heads; out.actors = actorIds;
}

// writes the header and the small sections (actor, actors, clock, deps) after the big record sections
inline void Engine::finishPatch(PatchOut& out) {
  if (out.bigEnd == 0) { out.propsOff = out.editsOff = PATCH_HDR_WORDS * 8; out.elemOff = 0; out.valBytesOff = out.valBytesLen = 0; out.bigEnd = PATCH_HDR_WORDS * 8; }
  size_t small = 64 + out.actor.size(); for (auto& a : out.actors) small += 8 + a.size(); small += out.clock.size() * 16 + out.deps.size() * 32 + 64;
  patchBuf.ensure(out.bigEnd + small);   // growth preserves what is already there
  u8* b = patchBuf.p; size_t at = out.bigEnd;
  auto pad8 = [&]() { while (at % 8) b[at++] = 0; };
  u64 hdr[PATCH_HDR_WORDS] = {0}; hdr[0] = 0x31504747414d41ULL; hdr[1] = out.maxOp; hdr[2] = out.pendingChanges; hdr[3] = out.hasActorSeq ? 1 : 0; hdr[4] = out.seq;
  pad8(); hdr[5] = at; hdr[6] = out.actor.size(); memcpy(b + at, out.actor.data(), out.actor.size()); at += out.actor.size(); pad8();
  hdr[7] = at; hdr[8] = out.actors.size();
  for (auto& a : out.actors) { const u32 l = (u32)a.size(); memcpy(b + at, &l, 4); at += 4; memcpy(b + at, a.data(), l); at += l; while (at % 4) b[at++] = 0; }
  pad8(); hdr[9] = at; hdr[10] = out.clock.size();
  for (auto& c : out.clock) { const u64 a = c.first, s = c.second; memcpy(b + at, &a, 8); memcpy(b + at + 8, &s, 8); at += 16; }
  hdr[11] = at; hdr[12] = out.deps.size(); for (auto& d : out.deps) { memcpy(b + at, d.data(), 32); at += 32; }
  hdr[13] = out.propsOff; hdr[14] = out.numProps; hdr[15] = out.editsOff; hdr[16] = out.numEdits; hdr[17] = out.elemOff; hdr[18] = out.valBytesOff; hdr[19] = out.valBytesLen;
  memcpy(b, hdr, sizeof(hdr));
  out.bytes = b; out.bytesLen = at;
}

// forgets the document but keeps every allocation (steady-state serving / benchmarking)
inline void Engine::reset() {
  sync(ctx); headIndexesUnknown = false; unknownCols.clear();
  arenaLen = 0; hostArena.len = 0; numApplied = 0; numRows = 0; numSucc = 0; dev_memset(ctx, succOff.p, 0, 4);
  actorIds.clear(); actorRep.clear(); clock.clear(); heads.clear(); headIdx.clear(); changes.clear(); changeHashes.clear(); deflatedOriginal.clear(); loadedDoc.clear(); numLoaded = 0; historyRebuilt = 0; haveHashGraph = true;
  queue.clear(); queueOriginal.clear(); maxOp = 0; rebuildActorTable();
}

// After Backend.load the hashes of the loaded changes are unknown until computeHashGraph has run. Like the reference
// (new.js:1833-1840) the first attempt goes without them; if a change then stays unapplied (or looks out of sequence)
// because it refers to history, the hash graph is computed and the call starts over. Nothing was committed by then.
struct NeedHistory {};
inline void Engine::applyChanges(const u8* const* bufs, const size_t* lens, size_t n, const u8* blob, const u64* offsets, bool isLocal, bool wantPatch, PatchOut& out, bool hostScan) {
  if (haveHashGraph) { applyChangesOnce(bufs, lens, n, blob, offsets, isLocal, wantPatch, out, hostScan); return; }
  bool retry = false;
  try { applyChangesOnce(bufs, lens, n, blob, offsets, isLocal, wantPatch, out, hostScan); }
  catch (NeedHistory&) { retry = true; }
  catch (Error& e) { if (e.code != AMG_ERR_RANGE) throw; retry = true; }
  if (!retry) return;
  drop_peeks(ctx);
  computeHashGraph();
  out = PatchOut();
  applyChangesOnce(bufs, lens, n, blob, offsets, isLocal, wantPatch, out, hostScan);
}
inline void Engine::applyChangesOnce(const u8* const* bufs, const size_t* lens, size_t n, const u8* blob, const u64* offsets, bool isLocal, bool wantPatch, PatchOut& out, bool hostScan) {
  PhaseTimer timer(ctx); HostClock hclk; int hmark = 12;
  NvtxPhases nvtx; nvtx.next("upload+hash+decode");   // NVTX ranges per pipeline phase (nsys / ncu timelines)
  for (auto& x : lastPhaseMs) x = 0;
  auto hostMark = [&]() { if (hmark < 24) lastPhaseMs[hmark++] = hclk.ms(); };
  curTimer = &timer; curHostMark = hostMark;
  static const bool liveMarks = getenv("AMG_DEBUG_LIVE") != nullptr;   // development aid: marks go to stderr as they happen (to see where a call is stuck)
  dbgMarks.clear(); dbgMark = [this, &hclk](const char* l) { dbgMarks.emplace_back(l, hclk.ms()); if (liveMarks) { fprintf(stderr, "amgpu mark %-28s %9.3f ms\n", l, hclk.ms()); fflush(stderr); } };
  struct SideJoinAll { Ctx& c; ~SideJoinAll() { side_join(c); } } sideJoinAll{ctx};   // whatever this call put on the side stream is ordered before the next call
  struct ClearTimer { Engine* e; ~ClearTimer() { e->curTimer = nullptr; e->curHostMark = nullptr; e->dbgMark = [](const char*) {}; } } clearTimer{this};
  // ------------------------------------------------------------ 0. stage the batch in the arena; hash and decode it piece by piece
  // The change bytes go to the device in pieces on the copy stream; as soon as a piece has landed, its changes are hashed
  // (side stream) and decoded (main stream) while the next piece is still crossing PCIe - both kernels read the piece
  // while it is hot in L2. The host keeps NO copy of bytes that came from a pinned or device buffer of the caller: the
  // mirror (hostArena) is filled lazily when something asks for it (getChanges, amg_arena ...; ensureHostMirror). Bytes
  // that have to be staged through pinned memory anyway (pageable caller buffers, pointer arrays) are staged through the
  // mirror itself, which then stays complete for free.
  const size_t arenaLen0 = arenaLen; const size_t hostLen0 = hostArena.size();
  std::vector<HostChange>& batch = batchStore; batch.clear();   // member: the 8 MB of a 1M-change batch keep their pages across calls
  std::vector<HostChange> batchOriginal, inflOrig;   // originals of DEFLATEd changes: batchOriginal (dense, queue entries) / inflOrig (sparse, parallel to deflIdx)
  std::vector<u32> deflIdx;
  size_t inflNd = 0, inflExtraStart = 0, inflExtra = 0; bool inflPending = false;
  // Host side of the device inflate: which batch entries moved where. Not on the critical path: runs when the information
  // is first needed (queue hand-over, commit).
  auto finishInflate = [&]() {
    if (!inflPending) return;
    inflPending = false; const size_t nd = inflNd;
    u32* origOff = patchTriples.p; u32* origLen = patchTriples.p + nd;
    pinnedScratch.ensure(5 * nd + 16); u32* ps = pinnedScratch.p;   // pinned: the five small copies queue up and complete with one sync
    d2h(ctx, ps, deflList.p, nd * 4); d2h(ctx, ps + nd, inflLen.p, nd * 4); d2h(ctx, ps + 2 * nd, inflOff.p, nd * 4);
    d2h(ctx, ps + 3 * nd, origOff, nd * 4); d2h(ctx, ps + 4 * nd, origLen, nd * 4);
    if (hostArena.size() == inflExtraStart) {   // the mirror is complete up to here: keep it complete
      hostArena.resize(inflExtraStart + inflExtra);
      d2h(ctx, hostArena.data() + inflExtraStart, arena.p + inflExtraStart, inflExtra);
    }
    sync(ctx);
    deflIdx.assign(ps, ps + nd); inflOrig.resize(nd);   // deflIdx is ascending: (batch index, original range), looked up by binary search
    for (size_t k = 0; k < nd; k++) { const u32 bi = ps[k]; inflOrig[k] = HostChange{ps[3 * nd + k], ps[4 * nd + k]}; batch[bi] = HostChange{(u32)inflExtraStart + ps[2 * nd + k], ps[nd + k]}; }
  };
  auto originalOf = [&](size_t b) -> HostChange {
    if (!batchOriginal.empty() && batchOriginal[b].len) return batchOriginal[b];
    auto it = std::lower_bound(deflIdx.begin(), deflIdx.end(), (u32)b);
    return it != deflIdx.end() && *it == (u32)b ? inflOrig[it - deflIdx.begin()] : HostChange{0, 0};
  };
  struct Rollback { Engine* e; size_t len; bool armed = true; ~Rollback() { if (armed) { if (e->hostArena.size() > len) e->hostArena.resize(len); e->rebuildActorTable(); } } };
  size_t total = 0;
  if (blob && n > 0) total = offsets[n] - offsets[0]; else for (size_t i = 0; i < n; i++) total += lens[i];
  if ((u64)arenaLen0 + total + 64 >= 0xfff00000ULL) throw Error(AMG_ERR_UNSUPPORTED, "amgpu: change arena limited to 4 GiB per document");
  Rollback rb{this, hostLen0};
  struct CopyJoin { Engine* e; ~CopyJoin() { copy_join(e->ctx); } } copyJoin{this};   // nothing of this call is left on the copy stream, also on the error paths
  size_t cur = arenaLen0;
  const size_t Bq = queue.size(); const size_t B = n + Bq;
  if (B == 0) { rb.armed = false; fillPatchHeader(out); finishPatch(out); return; }
  enum { SRC_PINNED, SRC_DEVICE, SRC_PAGEABLE } srcKind = SRC_PAGEABLE;
#ifndef AMG_EMU
  if (blob && n > 0) { cudaPointerAttributes at; if (cudaPointerGetAttributes(&at, blob) == cudaSuccess) { if (at.type == cudaMemoryTypeHost) srcKind = SRC_PINNED; else if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) srcKind = SRC_DEVICE; } else cudaGetLastError(); }
#endif
  const bool throughMirror = srcKind == SRC_PAGEABLE && total > 0;
  if (throughMirror) { ensureHostMirror(); hostArena.resize(arenaLen0 + total); }
  arena.ensure(ctx, arenaLen0 + total + 64, arenaLen0);
  dbgMark("stage:begin");
  batch.clear();
  // Pieces end on change boundaries. The copies of a pinned / device buffer are queued first (they need nothing but byte
  // ranges); the (offset, length) table of the changes is built and uploaded while they run; then every piece's kernels
  // are queued behind its copy. Pageable input is staged piece by piece, each piece's kernels right behind it.
  const size_t kPiece = 16u << 20;
  struct Piece { size_t byteEnd, changeEnd, mark; };
  std::vector<Piece> pieces;
  pairStage.ensure(B + 1); HostChange* pairs = pairStage.p;   // pinned: the table goes up by DMA while the host carries on
  if (blob && n > 0) {
    const size_t base = offsets[0];
    // cut points: every kPiece bytes, the tail halved three more times - what is left to hash and decode once the last byte
    // has arrived is a piece of a few MB, not a whole one
    std::vector<size_t> cuts; size_t o = kPiece;
    for (; o < total && total - o > kPiece; o += kPiece) cuts.push_back(o);
    if (total > 0) { const size_t from = o - kPiece; size_t rest = total - from; for (int k = 0; k < 3 && rest > (2u << 20); k++) { rest /= 2; cuts.push_back(total - rest); } }
    size_t lastCe = 0;
    for (size_t c : cuts) {
      const size_t ce = (size_t)(std::upper_bound(offsets, offsets + n + 1, (u64)(base + c)) - offsets) - 1;   // changes that end inside the first c bytes
      if (ce <= lastCe || ce >= n) continue;
      pieces.push_back({(size_t)(offsets[ce] - base), ce, 0}); lastCe = ce;
    }
    pieces.push_back({total, n, 0});
  } else {
    size_t at = 0, nextCut = kPiece;
    for (size_t i = 0; i < n; i++) {
      pairs[i] = HostChange{(u32)(arenaLen0 + at), (u32)lens[i]}; at += lens[i];
      if (at >= nextCut && i + 1 < n) { pieces.push_back({at, i + 1, 0}); nextCut = at + kPiece; }
    }
    pieces.push_back({at, n, 0});
  }
  cur = arenaLen0 + total;
  copy_fork(ctx);   // the copy stream starts behind what is queued on the main stream so far (arena growth)
  const bool copiesFirst = srcKind != SRC_PAGEABLE;
  auto queueCopy = [&](Piece& pc, size_t byte0, size_t ch0) {
    const size_t m = pc.byteEnd - byte0;
    if (m > 0) {
      u8* dst = arena.p + arenaLen0 + byte0;
      if (srcKind == SRC_PINNED) h2d_copy(ctx, dst, blob + offsets[0] + byte0, m);
      else if (srcKind == SRC_DEVICE) d2d_copy(ctx, dst, blob + offsets[0] + byte0, m);
      else {
        u8* stage = hostArena.data() + arenaLen0 + byte0;
        if (blob) parallel_copy(stage, blob + offsets[0] + byte0, m);
        else parallel_gather(stage, bufs, lens, ch0, pc.changeEnd);
        h2d_copy(ctx, dst, stage, m);
      }
    }
    pc.mark = copy_piece_record(ctx);
  };
  if (copiesFirst) { size_t byte0 = 0, ch0 = 0; for (Piece& pc : pieces) { queueCopy(pc, byte0, ch0); byte0 = pc.byteEnd; ch0 = pc.changeEnd; } }
  dbgMark("stage:copies-queued");
  // The (offset, length) table of the changes. Packed batch whose offsets array is pinned or device memory: the array goes
  // up by DMA and a kernel derives the table (the host's own copy is filled later, in the shadow of the device work).
  // Otherwise the host fills a pinned table (a few threads for 1M entries) and uploads that.
  auto fillPairs = [&]() {
    if (!(blob && n > 0)) return;
    const size_t base = offsets[0]; const u32 shift = (u32)(arenaLen0 - base);
    auto fill = [&](size_t a, size_t b) { for (size_t i = a; i < b; i++) { pairs[i].off = (u32)offsets[i] + shift; pairs[i].len = (u32)(offsets[i + 1] - offsets[i]); } };
    const unsigned nt = n < (1u << 16) ? 1u : std::min<unsigned>(4, std::max(1u, std::thread::hardware_concurrency()));
    if (nt == 1) fill(0, n);
    else { std::vector<std::thread> ts; const size_t per = (n + nt - 1) / nt; for (unsigned t = 0; t < nt; t++) { const size_t a = t * per, b = std::min(n, a + per); if (a < b) ts.emplace_back(fill, a, b); } for (auto& t : ts) t.join(); }
  };
  bool offsetsByDma = false;
#ifndef AMG_EMU
  if (blob && n >= 4096) { cudaPointerAttributes at; if (cudaPointerGetAttributes(&at, offsets) == cudaSuccess) offsetsByDma = at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged; else cudaGetLastError(); }
#endif
  for (size_t i = 0; i < Bq; i++) pairs[n + i] = queue[i];
  chPairs.ensure(ctx, B); chOff.ensure(ctx, B); chLen.ensure(ctx, B);
  if (offsetsByDma) {
    DBuf<u64>& offsD = offsDev; offsD.ensure(ctx, n + 2);
    CUDA_CHECK_EMU(cudaMemcpyAsync(offsD.p, offsets, (n + 1) * 8, cudaMemcpyDefault, ctx.stream));
    foreach(ctx, n, OffsetsToRangesKernel{offsD.p, (u32)(arenaLen0 - offsets[0]), chOff.p, chLen.p});
    if (Bq > 0) { h2d(ctx, chPairs.p + n, pairs + n, Bq * sizeof(HostChange)); foreach(ctx, Bq, SplitPairsKernel{chPairs.p + n, chOff.p + n, chLen.p + n}); }
  } else {
    fillPairs();
    h2d(ctx, chPairs.p, pairs, B * sizeof(HostChange));
    foreach(ctx, B, SplitPairsKernel{chPairs.p, chOff.p, chLen.p});
  }
  dev_memset(ctx, arena.p + cur, 0, 64);
  dev_memset(ctx, errWord.p, 0, 16); errSnapLaunches = ~0ull;
  hashes.ensure(ctx, (numApplied + B) * 32 + 64, numApplied * 32);
  deflList.ensure(ctx, B + 1);
  u8* hashOut = hashes.p + numApplied * 32;
  const DecodeTilesArgs dargs = decodeArgs(arena.p, B, cur - arenaLen0);
  decode_tiles_begin(ctx, dargs);
  struct SideJoin { Ctx& c; ~SideJoin() { side_join(c); } } sideJoin{ctx};   // also on the error paths: nothing of this call outlives it on the side stream
  dbgMark("stage:tables");
  // changes [c0, c1) are on the device once the copy stream has passed the piece's mark: hash on the side stream, decode on the main one
  auto processRange = [&](size_t c0, size_t c1, bool waitCopy, size_t mark) {
    if (c1 <= c0) return;
    if (waitCopy) copy_piece_wait(ctx, mark);   // both streams wait for the piece
    else side_fork(ctx);
    sha_range(ctx, ShaTilesArgs{arena.p, chOff.p, chLen.p, hashOut, errWord.p, deflList.p, (u32)c0, (u32)c1}, true);
    decode_tiles_range(ctx, dargs, (u32)c0, (u32)c1);
  };
  side_fork(ctx);   // the side stream is ordered behind the tables
  if (Bq > 0) processRange(n, B, false, 0);   // queue entries: their bytes are on the device already
  {
    size_t byte0 = 0, ch0 = 0;
    for (Piece& pc : pieces) {
      if (!copiesFirst) queueCopy(pc, byte0, ch0);
      processRange(ch0, pc.changeEnd, true, pc.mark);
      byte0 = pc.byteEnd; ch0 = pc.changeEnd;
    }
  }
  dbgMark("stage:enqueued");
  timer.mark(); hostMark();
  // The host's own list of the batch entries (bookkeeping at commit, queue hand-over) is filled by a helper thread while
  // this one keeps the device fed; needBatch() joins it before the list is first read.
  struct BatchFill { std::thread t; void join() { if (t.joinable()) t.join(); } ~BatchFill() { join(); } } batchFill;
  auto fillBatch = [&, pairs, B, n, Bq]() {
    if (offsetsByDma) fillPairs();
    batch.assign(pairs, pairs + B);
    if (Bq > 0) { batchOriginal.assign(B, HostChange{0, 0}); for (size_t i = 0; i < Bq; i++) batchOriginal[n + i] = queueOriginal[i]; }
  };
  if (B >= (1u << 15)) batchFill.t = std::thread(fillBatch); else fillBatch();
  auto needBatch = [&]() { batchFill.join(); };
  nvtx.next("inflate+decode-finish");
  // ------------------------------------------------------------ 1. DEFLATEd changes
  {
    // Which changes of the batch are DEFLATEd (columnar.js:742)? Those are inflated on the device, behind the batch:
    // flag -> scan -> ordered list -> k_inflate (decode into scratch, sizes) -> scan -> k_inflate (assemble in place); the originals stay.
    // The hash / decode kernels above skipped them; they are hashed and decoded here, from the inflated bytes.
    emit.ensure(ctx, B + 1); slot.ensure(ctx, B + 2);
    dev_memset(ctx, flagWord.p + 8, 0, 4);
    foreach(ctx, B, DeflateFlagKernel{arena.p, chOff.p, chLen.p, emit.p, flagWord.p + 8});
    scan_exclusive(ctx, scanTmp, emit.p, slot.p, B);
    u32 nd32 = 0, deflBytes = 0; readU32x2(slot.p + B, flagWord.p + 8, &nd32, &deflBytes);
    const size_t nd = nd32;
    dbgMark("sha:deflate-scanned");
    if (nd > 0) {
      // every stream is decoded once, into scratch (capacity: a few times its compressed size); the sizes give the places
      // behind the batch, a second kernel assembles the changes there (copy; the rare stream that did not fit is decoded again)
      foreach(ctx, B, CompactKernel{emit.p, slot.p, deflList.p});
      inflLen.ensure(ctx, nd + 1); inflOff.ensure(ctx, nd + 2); patchTriples.ensure(ctx, 2 * nd + 2); inflCap.ensure(ctx, nd + 2); inflCapOff.ensure(ctx, nd + 2); inflOvf.ensure(ctx, nd + 1);
      u32* origOff = patchTriples.p; u32* origLen = patchTriples.p + nd;
      u32 factor = 4; while (factor > 1 && (u64)factor * deflBytes + 1024ull * nd >= 0xf0000000ULL) factor--;
      const size_t scratchBytes = (size_t)factor * deflBytes + 1024 * nd + 64;
      inflScratch.ensure(ctx, scratchBytes);
      foreach(ctx, nd, InflateCapKernel{deflList.p, chLen.p, factor, inflCap.p});
      scan_exclusive(ctx, scanTmp, inflCap.p, inflCapOff.p, nd);
      InflateArgs ia{arena.p, chOff.p, chLen.p, deflList.p, nd, inflLen.p, nullptr, 0u, origOff, origLen, inflScratch.p, inflCapOff.p, inflOvf.p, errWord.p};
      inflate_changes(ctx, INFL_SPECULATE, ia);
      scan_exclusive(ctx, scanTmp, inflLen.p, inflOff.p, nd);
      const size_t extra = readU32(inflOff.p + nd);
      if (errSnapshot) throwKernelError(errSnapshot, actorIds);   // (the error word travels with every small read)
      if ((u64)cur + extra + 64 >= 0xfff00000ULL) throw Error(AMG_ERR_UNSUPPORTED, "amgpu: change arena limited to 4 GiB per document");
      const size_t extraStart = cur; cur += extra;
      side_join(ctx);   // the arena may move: nothing may still be reading it
      arena.ensure(ctx, cur + 64, extraStart);
      ia.arena = arena.p; ia.outOff = inflOff.p; ia.extraStart = (u32)extraStart;
      inflate_changes(ctx, INFL_PLACE, ia);
      dev_memset(ctx, arena.p + cur, 0, 64);
      foreach(ctx, nd, InflatePatchKernel{deflList.p, inflLen.p, inflOff.p, (u32)extraStart, chOff.p, chLen.p});
      foreach(ctx, nd, ShaKernel{arena.p, chOff.p, chLen.p, hashOut, errWord.p, deflList.p, nullptr});
      inflNd = nd; inflExtraStart = extraStart; inflExtra = extra; inflPending = true;   // host bookkeeping happens in finishInflate()
      dbgMark("sha:inflated");
    }
  }
  // changes the tile kernel passed on (inflated ones, changes outside their tile's window), then the totals
  { DecodeTilesArgs fin = dargs; fin.arena = arena.p; decode_tiles_list(ctx, fin, deflList.p, (u32)inflNd); decode_tiles_finish(ctx, fin, B); }
  lastDeflCount = inflNd; lastDeflStart = inflExtraStart;
  dbgMark("decode:finish-enqueued");
  side_join(ctx);
  timer.mark(); hostMark(); nvtx.next("gate");
  // (parse errors surface with the first host round trip of the gate: the error word travels with every small read, and a
  //  change that failed to parse has zero deps / ops so the kernels in between have nothing to walk)
  // ------------------------------------------------------------ 2. causal gate
  depBase.ensure(ctx, B + 1); scan_exclusive(ctx, scanTmp, nDeps.p, depBase.p, B);
  const size_t depBound = (cur - arenaLen0) / 32 + B + 1;   // every dependency occupies 32 bytes of its change: no need to read the exact total
  depIdx.ensure(ctx, depBound + 1); primary.ensure(ctx, B); pass.ensure(ctx, B);
  const size_t G = numApplied + B; const size_t tcap = pow2_at_least(2 * G + 2);
  hashTable.ensure(ctx, tcap); dev_memset(ctx, hashTable.p, 0xff, tcap * 4);
  foreach(ctx, G, HashInsertKernel{hashes.p, hashTable.p, (u64)tcap - 1});
  foreach(ctx, B, ResolveDepsKernel{arena.p, hashes.p, hashTable.p, (u64)tcap - 1, hot.p, nDeps.p, numApplied, depBase.p, depIdx.p, primary.p});
  fill32(pass.p, 1, B);
  dev_memset(ctx, flagWord.p + 12, 0, 4);
  foreach(ctx, B, GateDupFlagKernel{primary.p, numApplied, flagWord.p + 12});
  bool copiesChecked = false, haveCopies = false; u32 decTot[4] = {0, 0, 0, 0};
  for (size_t iter = 0; iter <= B + 1; iter += 2) {   // two sweeps per host round trip: the common batch settles in the first
    if (haveCopies) break;
    foreach(ctx, B, RelaxKernel{depBase.p, depIdx.p, nDeps.p, primary.p, numApplied, pass.p, flagWord.p, (u32)B + 1});
    dev_memset(ctx, flagWord.p, 0, 4);
    foreach(ctx, B, RelaxKernel{depBase.p, depIdx.p, nDeps.p, primary.p, numApplied, pass.p, flagWord.p, (u32)B + 1});
    u32 again = 0, copies = 0;
    { void* dst[6] = {&again, &copies, &decTot[0], &decTot[1], &decTot[2], &decTot[3]};
      readWords({{flagWord.p, 4}, {flagWord.p + 12, 4}, {decTotalsPtr(), 4}, {decTotalsPtr() + 1, 4}, {decTotalsPtr() + 2, 4}, {decTotalsPtr() + 3, 4}}, dst); }
    checkErr(actorIds);   // free: the error word came with the read
    if (iter == 0 && decodeOverflowed(decTot)) {   // the raw row tables were too small for this batch: grown, decoded again (same results otherwise)
      runDecodeTiles(arena.p, B, cur - arenaLen0, deflList.p, inflNd, inflExtraStart);   // the whole batch is resident by now (inflated changes re-pointed)
      void* d2[4] = {&decTot[0], &decTot[1], &decTot[2], &decTot[3]};
      readWords({{decTotalsPtr(), 4}, {decTotalsPtr() + 1, 4}, {decTotalsPtr() + 2, 4}, {decTotalsPtr() + 3, 4}}, d2);
      if (decTot[2]) throw Error(AMG_ERR_INTERNAL, "amgpu: decode row tables overflowed twice");
    }
    if (!copiesChecked) { copiesChecked = true; haveCopies = copies != 0; }
    if (!again) break;
  }
  if (haveCopies) {   // rare: a change that is waiting was delivered again (gate.cuh, GateBestKernel ...)
    gateBest.ensure(ctx, B + 1); fill32(pass.p, 1, B);
    for (size_t iter = 0; iter <= B + 1; iter++) {
      dev_memset(ctx, gateBest.p, 0xff, B * 8); dev_memset(ctx, flagWord.p, 0, 4);
      foreach(ctx, B, GateBestKernel{primary.p, pass.p, numApplied, gateBest.p});
      foreach(ctx, B, RelaxCopiesKernel{depBase.p, depIdx.p, nDeps.p, primary.p, numApplied, gateBest.p, pass.p, flagWord.p, (u32)B + 1});
      const u32 again = readU32(flagWord.p);
      checkErr(actorIds);
      if (!again) break;
    }
    dev_memset(ctx, gateBest.p, 0xff, B * 8);
    foreach(ctx, B, GateBestKernel{primary.p, pass.p, numApplied, gateBest.p});
    const size_t depTotal = readU32(depBase.p + B);
    if (depTotal) foreach(ctx, depTotal, GateDepWinnerKernel{gateBest.p, numApplied, depIdx.p});
    foreach(ctx, B, GateWinnerKernel{gateBest.p, numApplied, primary.p});
  }
  dbgMark("gate:settled");
  applied.ensure(ctx, B); appRank.ensure(ctx, B + 1); isRow.ensure(ctx, B + 1);
  dev_memset(ctx, flagWord.p, 0, 8);
  foreach(ctx, B, AppliedFlagKernel{primary.p, pass.p, numApplied, applied.p, isRow.p, flagWord.p});
  u32 stats[2]; readU32x2(flagWord.p, flagWord.p + 1, &stats[0], &stats[1]);
  const size_t numNew = stats[0]; const bool inOrder = stats[1] <= 1; batchInOrder = inOrder;
  std::vector<u8> appliedH; std::vector<u32> primaryH, appRankH;
  if (inOrder) scan_exclusive(ctx, scanTmp, isRow.p, appRank.p, B);
  else {
    sortKeys.ensure(ctx, B); sortVals.ensure(ctx, B);
    foreach(ctx, B, PassKeyKernel{pass.p, applied.p, sortKeys.p, sortVals.p});
    sortPairs(sortKeys, sortVals, B, 32);
    foreach(ctx, B, RankFromOrderKernel{sortVals.p, appRank.p, numNew});
  }
  if (numNew < B || !inOrder) {
    appliedH.resize(B); primaryH.resize(B); appRankH.resize(B);
    d2h(ctx, appliedH.data(), applied.p, B); d2h(ctx, primaryH.data(), primary.p, B * 4); d2h(ctx, appRankH.data(), appRank.p, B * 4); sync(ctx);
  }
  if (!haveHashGraph && numNew < B) throw NeedHistory{};   // a change waits for (or repeats) something older than the loaded heads
  // the queue after this call: every batch entry whose hash is still not applied (new.js:1569-1570, 1832)
  std::vector<HostChange> newQueue, newQueueOriginal;
  if (numNew < B) { needBatch(); finishInflate(); }
  if (numNew < B) for (size_t b = 0; b < B; b++) {
    const u32 pr = primaryH[b];
    const bool hashApplied = pr < numApplied || appliedH[pr - numApplied];
    if (!hashApplied) { newQueue.push_back(batch[b]); newQueueOriginal.push_back(originalOf(b)); }
  }
  timer.mark(); hostMark(); nvtx.next("actors+seq+finalize");
  std::vector<std::string> actorsNow = actorIds; std::vector<u64> clockNow = clock; std::vector<u32> actorCntH; std::vector<std::pair<u32, u32>> actorRepNow = actorRep;
  size_t M = 0, P = 0, N = numRows, numPairs = numSucc; u64 maxOpNow = maxOp; bool hasUnknownColsCall = false;
  IdTable idt{nullptr, nullptr, 0};
  std::vector<std::array<u8, 32>> headsNow = heads; std::vector<u32> headIdxNow;
  if (numNew > 0) {
    // ---------------------------------------------------------- 3. actors
    authorSlot.ensure(ctx, B); newSlots.ensure(ctx, B + 1); u32 fresh = 0;
    while (true) {   // grow the table until the distinct authors fit at load factor <= 1/2
      dev_memset(ctx, flagWord.p, 0, 8);
      foreach(ctx, B, ActorInternKernel{arena.p, hot.p, applied.p, appRank.p, actorSlots.p, (u64)actorCap - 1, authorSlot.p, flagWord.p + 1});
      foreach(ctx, B, NewActorKernel{hot.p, applied.p, authorSlot.p, actorSlots.p, newSlots.p, flagWord.p});
      u32 full = 0; readU32x2(flagWord.p, flagWord.p + 1, &fresh, &full);
      if (!full && (actorIds.size() + fresh) * 2 <= actorCap) break;
      actorCap *= 4; actorSlots.ensure(ctx, actorCap); rebuildActorTable();
    }
    if (fresh > 0) {
    dbgMark("actors:interned");
      // slot numbers, slot records and id bytes of the new actors in ONE round trip (gathered into a staging buffer)
      static const u32 ACTOR_STAGE = 64;
      hashTmp.ensure(ctx, (size_t)fresh * (4 + sizeof(ActorSlot) + ACTOR_STAGE) + 64);
      u8* stageD = hashTmp.p; const size_t recsAt = ((size_t)fresh * 4 + 15) & ~(size_t)15, bytesAt = recsAt + (size_t)fresh * sizeof(ActorSlot), stageLen = bytesAt + (size_t)fresh * ACTOR_STAGE;
      foreach(ctx, fresh, GatherNewActorsKernel{arena.p, actorSlots.p, newSlots.p, reinterpret_cast<u32*>(stageD), reinterpret_cast<ActorSlot*>(stageD + recsAt), stageD + bytesAt, ACTOR_STAGE});
      std::vector<u8> stageH(stageLen); d2h(ctx, stageH.data(), stageD, stageLen); sync(ctx);
      std::vector<u32> slotsH(fresh); memcpy(slotsH.data(), stageH.data(), (size_t)fresh * 4);
      std::vector<ActorSlot> recs(fresh); memcpy(recs.data(), stageH.data() + recsAt, (size_t)fresh * sizeof(ActorSlot));
      std::vector<u32> order(fresh); for (u32 i = 0; i < fresh; i++) order[i] = i;
      std::sort(order.begin(), order.end(), [&](u32 a, u32 b) { return recs[a].first < recs[b].first; });
      std::vector<u32> ids(fresh), nums(fresh); actorsNow.reserve(actorsNow.size() + fresh);   // async copies target the strings: no reallocation below
      bool longIds = false;
      for (u32 k = 0; k < fresh; k++) {
        const ActorSlot& r = recs[order[k]]; ids[k] = slotsH[order[k]]; nums[k] = (u32)actorsNow.size();
        actorsNow.emplace_back(r.repLen, '\0');
        if (r.repLen <= ACTOR_STAGE) memcpy(&actorsNow.back()[0], stageH.data() + bytesAt + (size_t)order[k] * ACTOR_STAGE, r.repLen);
        else { d2h(ctx, &actorsNow.back()[0], arena.p + r.repOff, r.repLen); longIds = true; }   // from the device copy: the host mirror may still be filling
        actorRepNow.emplace_back(r.repOff, r.repLen);
      }
      if (longIds) sync(ctx);
      if (actorsNow.size() > 65535) throw Error(AMG_ERR_UNSUPPORTED, "amgpu: more than 65535 actors in one document");
      sortVals.ensure(ctx, 2 * fresh); h2d(ctx, sortVals.p, ids.data(), fresh * 4); h2d(ctx, sortVals.p + fresh, nums.data(), fresh * 4);
      foreach(ctx, fresh, SetActorNumKernel{actorSlots.p, sortVals.p, sortVals.p + fresh});
    }
    const size_t A = actorsNow.size(); clockNow.resize(A, 0);
    {   // rank of every actor in hex-string order (== byte order of the raw ids; new.js:64-65, 1180, 1198)
      std::vector<u32> order(A), rank(A); for (size_t i = 0; i < A; i++) order[i] = (u32)i;
      std::sort(order.begin(), order.end(), [&](u32 a, u32 b) { return actorsNow[a] < actorsNow[b]; });
      for (size_t i = 0; i < A; i++) rank[order[i]] = (u32)i;
      actorRank.ensure(ctx, A + 1); h2d(ctx, actorRank.p, rank.data(), A * 4);
    }
    const int rb = bits_for(A > 1 ? A - 1 : 1);
    dbgMark("actors:numbered");
    Ord ord{actorRank.p, rb};
    amapBase.ensure(ctx, B + 1); rowSlot.ensure(ctx, B + 1);
    foreach(ctx, B, MaskedCountKernel{nActors.p, applied.p, rowSlot.p});
    scan_exclusive(ctx, scanTmp, rowSlot.p, amapBase.p, B);
    amap.ensure(ctx, B + (cur - arenaLen0) / 2 + 2);   // author + one entry per other-actor table entry (>= 2 bytes each): bound instead of a round trip
    foreach(ctx, B, ActorMapKernel{arena.p, hot.p, nActors.p, applied.p, appRank.p, actorSlots.p, (u64)actorCap - 1, amapBase.p, amap.p, errWord.p});
    dbgMark("actors:mapped");
    // ---------------------------------------------------------- 4. sequence numbers
    changeActor.ensure(ctx, B); actorCnt.ensure(ctx, A + 1); actorBaseD.ensure(ctx, A + 1); seqSlot.ensure(ctx, numNew + 1);
    dev_memset(ctx, actorCnt.p, 0, (A + 1) * 4);
    foreach(ctx, B, ChangeActorKernel{amapBase.p, amap.p, applied.p, changeActor.p, actorCnt.p});
    {
      const u64 ew = fetchErr();
      if ((ew & 0xff) == KE_UNKNOWN_ACTOR) {   // name the actor like the reference does (new.js:1446): re-read that change's actor table
        const size_t b = (size_t)(ew >> 8); ChangeHot m0; u32 na0 = 1; d2h(ctx, &m0, hot.p + b, sizeof(ChangeHot)); d2h(ctx, &na0, nActors.p + b, 4); sync(ctx);
        std::vector<u8> bytes(m0.len); d2h(ctx, bytes.data(), arena.p + m0.off, m0.len); sync(ctx);
        ByteReader r(bytes.data(), m0.otherOff - m0.off, m0.len); std::string culprit;
        for (u32 k = 0; k < na0 && !r.err; k++) {
          u32 off, len; if (k == 0) { off = m0.actorOff - m0.off; len = m0.actorLen; } else { len = (u32)r.uleb(); off = r.pos; r.skip(len); }
          if (r.err) break;
          const std::string id((const char*)bytes.data() + off, len);
          if (std::find(actorsNow.begin(), actorsNow.end(), id) == actorsNow.end()) { culprit = id; break; }
        }
        if (!culprit.empty()) throw Error(AMG_ERR_RANGE, "actorId " + hex_of((const u8*)culprit.data(), culprit.size()) + " is not known to document");
      }
      if (ew) throwKernelError(ew, actorsNow);
    }
    scan_exclusive(ctx, scanTmp, actorCnt.p, actorBaseD.p, A);
    DBuf<u64>& clockDev = pairKey;   // scratch reuse before the succ phase
    clockDev.ensure(ctx, A + 1); h2d(ctx, clockDev.p, clockNow.data(), A * 8);
    dev_memset(ctx, seqSlot.p, 0xff, (numNew + 1) * 4); dev_memset(ctx, flagWord.p, 0, 4);
    foreach(ctx, B, SeqScatterKernel{hot.p, applied.p, changeActor.p, appRank.p, actorBaseD.p, actorCnt.p, clockDev.p, seqSlot.p, flagWord.p});
    foreach(ctx, B, SeqMonoKernel{hot.p, applied.p, changeActor.p, actorBaseD.p, actorCnt.p, clockDev.p, seqSlot.p, flagWord.p});
    actorCntH.resize(A); d2h(ctx, actorCntH.data(), actorCnt.p, A * 4);
    if (readU32(flagWord.p)) {
      // error path: replay the sequence check in application order on the host to produce the reference's message
      std::vector<ChangeHot> mh(B); std::vector<u32> ca(B), ar(B); std::vector<u8> ap(B);
      d2h(ctx, mh.data(), hot.p, B * sizeof(ChangeHot)); d2h(ctx, ca.data(), changeActor.p, B * 4); d2h(ctx, ar.data(), appRank.p, B * 4); d2h(ctx, ap.data(), applied.p, B); sync(ctx);
      std::vector<u32> byRank(numNew, 0); for (size_t b = 0; b < B; b++) if (ap[b]) byRank[ar[b]] = (u32)b;
      std::vector<u64> clk = clockNow;
      for (size_t k = 0; k < numNew; k++) {
        const u32 b = byRank[k]; const u64 expected = clk[ca[b]] + 1; const std::string actorHex = hex_of((const u8*)actorsNow[ca[b]].data(), actorsNow[ca[b]].size());
        if (mh[b].seq < expected) throw Error(AMG_ERR_RANGE, "Reuse of sequence number " + std::to_string(mh[b].seq) + " for actor " + actorHex);
        if (mh[b].seq > expected) throw Error(AMG_ERR_RANGE, "Skipped sequence number " + std::to_string(expected) + " for actor " + actorHex);
        clk[ca[b]] = mh[b].seq;
      }
      throw Error(AMG_ERR_INTERNAL, "amgpu: sequence check disagreement");
    }
    for (size_t a = 0; a < A; a++) clockNow[a] += actorCntH[a];
    dbgMark("seq:checked");
    // ---------------------------------------------------------- 5. ops of the applied changes
    // The rows were decoded with the headers (step 1), in batch order. When every change of the batch is applied the
    // op / pred ranges of the changes are the raw ones; otherwise they are the scans over the applied changes only.
    timeBase.ensure(ctx, B + 1);
    const bool allApplied = numNew == B;
    opBase.ensure(ctx, B + 1); predBase.ensure(ctx, B + 1); u32* opBaseP = opBase.p; u32* predBaseP = predBase.p;
    const u32 anyLarge = decTot[3] & 1u; hasUnknownColsCall = (decTot[3] & 2u) != 0;
    // first op / pred of every applied change in batch order (the raw rows themselves lie in tile arrival order, rawBase)
    if (allApplied) { scan_exclusive(ctx, scanTmp, nOps.p, opBase.p, B); scan_exclusive(ctx, scanTmp, nPreds.p, predBase.p, B); M = decTot[0]; P = decTot[1]; }
    else {
      foreach(ctx, B, MaskedCountKernel{nOps.p, applied.p, rowSlot.p}); scan_exclusive(ctx, scanTmp, rowSlot.p, opBase.p, B);
      foreach(ctx, B, MaskedCountKernel{nPreds.p, applied.p, rowSlot.p}); scan_exclusive(ctx, scanTmp, rowSlot.p, predBase.p, B);
      u32 m32 = 0, p32 = 0; void* dst[2] = {&m32, &p32}; readWords({{opBase.p + B, 4}, {predBase.p + B, 4}}, dst); M = m32; P = p32;
    }
    if (decTot[0] >= 0x7fffffffu || decTot[1] >= 0x7fffffffu) throw Error(AMG_ERR_UNSUPPORTED, "amgpu: more than 2^31 operations in one call");
    dbgMark("decode:counts");
    if (!inOrder) {
      perm.ensure(ctx, B + 1); dev_memset(ctx, perm.p, 0, (B + 1) * 4);
      foreach(ctx, B, OpsInOrderKernel{nOps.p, applied.p, appRank.p, perm.p});
      scan_exclusive(ctx, scanTmp, perm.p, perm.p, numNew);
    }
    foreach(ctx, B, TimeBaseKernel{perm.p, applied.p, appRank.p, opBaseP, inOrder ? 1 : 0, timeBase.p});
    DBuf<u64>& maxOpD = pairSucc; maxOpD.ensure(ctx, 1); h2d(ctx, maxOpD.p, &maxOpNow, 8);
    foreach(ctx, B, MaxOpKernel{hot.p, nOps.p, applied.p, decErr.p, maxOpD.p, errWord.p});
    d2h(ctx, &maxOpNow, maxOpD.p, 8);
    RawRows raw = rawRows();
    lastNumLarge = 0;
    if (anyLarge) {   // changes with more than SMALL_CHANGE_OPS ops: (column, change)-parallel expansion into their reserved rows
      largeFlag.ensure(ctx, B + 1); largeSlot.ensure(ctx, B + 2); largeList.ensure(ctx, B + 1);
      foreach(ctx, B, LargeFlagKernel{nOps.p, applied.p, largeFlag.p});
      scan_exclusive(ctx, scanTmp, largeFlag.p, largeSlot.p, B);
      lastNumLarge = readU32(largeSlot.p + B);
      if (lastNumLarge > 0) {
        foreach(ctx, B, CompactKernel{largeFlag.p, largeSlot.p, largeList.p});
        // bulk changes (thousands of ops in one change): their columns are expanded in parallel by the token / record
        // decoders of doccols.cuh, column by column; whatever those decline (non-canonical streams, columns that do not hold
        // exactly the op count) and all other large changes go through DecodeColumnKernel (one thread per column)
        u32 hugeMask = 0;
        if (lastNumLarge <= 8) hugeMask = decodeHugeChanges(raw, lastNumLarge);
        foreach(ctx, (size_t)NCOLS * lastNumLarge, DecodeColumnKernel{arena.p, largeList.p, lastNumLarge, hot.p, nOps.p, nPreds.p, rawBase.p, rawPredBase.p, applied.p, raw, errWord.p, hugeDone.p});
        (void)hugeMask;
      }
    }
    for (DBuf<u64>* b : {&o_id, &o_obj, &o_key}) b->ensure(ctx, M + 1);
    o_predId.ensure(ctx, P + 1);
    for (DBuf<u32>* b : {&o_keyStrOff, &o_keyStrLen, &o_flags, &o_valLen, &o_valOff, &o_predOff, &o_predNum, &o_change, &o_time}) b->ensure(ctx, M + 1);
    OpRows ops{o_id.p, o_obj.p, o_key.p, o_keyStrOff.p, o_keyStrLen.p, o_flags.p, o_valLen.p, o_valOff.p, o_predOff.p, o_predNum.p, o_change.p, o_time.p, o_predId.p};
    foreach(ctx, M, FinalizeOpsKernel{B, hot.p, nActors.p, opBaseP, predBaseP, rawBase.p, rawPredBase.p, timeBase.p, amapBase.p, amap.p, applied.p, raw, ops, errWord.p});
    checkErr(actorsNow);
    dbgMark("decode:finalized");
    timer.mark(); hostMark(); nvtx.next("opset");
    // ---------------------------------------------------------- 6. op set
    if (maxOpNow >= (1ULL << 40)) throw Error(AMG_ERR_UNSUPPORTED, "amgpu: op counters above 2^40");
    const int ordBits = bits_for(maxOpNow) + rb;
    isRow.ensure(ctx, M + 1); rowSlot.ensure(ctx, M + 1); rowOfOp.ensure(ctx, M + 1);
    foreach(ctx, M, RowFlagKernel{o_flags.p, isRow.p}); scan_exclusive(ctx, scanTmp, isRow.p, rowSlot.p, M);
    const size_t R = readU32(rowSlot.p + M); N = numRows + R;
    if (N >= (1u << 29)) throw Error(AMG_ERR_UNSUPPORTED, "amgpu: more than 2^29 document rows");
    doc.ensure(ctx, N + 1, numRows);
    DocRows w = doc.view();
    foreach(ctx, M, AppendRowsKernel{ops, isRow.p, rowSlot.p, numRows, w, rowOfOp.p});
    const size_t icap = pow2_at_least(2 * N + 2); idKeys.ensure(ctx, icap); idVals.ensure(ctx, icap); dev_memset(ctx, idKeys.p, 0, icap * 8);
    idt = IdTable{idKeys.p, idVals.p, (u64)icap - 1};
    foreach(ctx, N, IdInsertKernel{w.id, idt, errWord.p});
    objRow.ensure(ctx, N + 1); elemRow.ensure(ctx, N + 1); parentRow.ensure(ctx, N + 1);
    foreach(ctx, N, ResolveRowsKernel{w, idt, ord, numRows, objRow.p, elemRow.p, parentRow.p, errWord.p});
    checkErr(actorsNow);
    dbgMark("opset:resolved");
    // map keys: intern, verify, rank distinct keys with an LSD string sort
    const size_t kcap = pow2_at_least(2 * N + 2); keySlots.ensure(ctx, kcap); keySlot.ensure(ctx, N + 1); repList.ensure(ctx, N + 1); repCount.ensure(ctx, 4);
    foreach(ctx, kcap, KeySlotInitKernel{keySlots.p});
    foreach(ctx, N, KeyInternKernel{arena.p, w, keySlots.p, (u64)kcap - 1, keySlot.p});
    dev_memset(ctx, repCount.p, 0, 16);
    foreach(ctx, N, KeyVerifyKernel{arena.p, w, keySlots.p, keySlot.p, repList.p, repCount.p, errWord.p});
    u32 rc[2]; readU32x2(repCount.p, repCount.p + 1, &rc[0], &rc[1]);
    const size_t D = rc[0]; const u32 maxKeyLen = rc[1];
    dbgMark("opset:keys-interned");
    if (D > 0) {
      sortKeys.ensure(ctx, D); sortVals.ensure(ctx, D);
      d2d(ctx, sortVals.p, repList.p, D * 4);
      const int chunks = (int)((maxKeyLen + 6) / 7);
      for (int ch = std::max(chunks, 1) - 1; ch >= 0; ch--) {
        foreach(ctx, D, KeyChunkKernel{arena.p, w, sortVals.p, (u32)ch * 7, sortKeys.p});
        sortPairs(sortKeys, sortVals, D, 64);
      }
      foreach(ctx, D, KeyRankKernel{sortVals.p, keySlot.p, keySlots.p});
    }
    // RGA order of every list: sibling sort, Euler tour, pointer jumping
    listPos.ensure(ctx, N + 1); insItems.ensure(ctx, N + 1); emit.ensure(ctx, N + 1); slot.ensure(ctx, N + 2);
    foreach(ctx, N, InsertFlagKernel{w, emit.p}); scan_exclusive(ctx, scanTmp, emit.p, slot.p, N);
    const size_t I = readU32(slot.p + N);
    dbgMark("opset:inserts-counted");
    if (I > 0) {
      const int parentBits = bits_for(N) + 1;
      if (ordBits + parentBits > 64) throw Error(AMG_ERR_UNSUPPORTED, "amgpu: opId range x document size exceeds the 64-bit sibling sort key");
      foreach(ctx, N, CompactKernel{emit.p, slot.p, insItems.p});
      sortKeys.ensure(ctx, I); sortVals.ensure(ctx, I);
      foreach(ctx, I, SiblingKeyKernel{w, insItems.p, parentRow.p, objRow.p, ord, ordBits, sortKeys.p, errWord.p});
      d2d(ctx, sortVals.p, insItems.p, I * 4);
      sortPairs(sortKeys, sortVals, I, ordBits + parentBits);
      checkErr(actorsNow);
      itemIdx.ensure(ctx, N + 1); objSlot.ensure(ctx, N + 2);
      foreach(ctx, I, ItemIndexKernel{sortVals.p, itemIdx.p});
      foreach(ctx, N, ListObjFlagKernel{w, emit.p}); scan_exclusive(ctx, scanTmp, emit.p, objSlot.p, N);
      const size_t Lo = readU32(objSlot.p + N);
      const size_t S = 2 * I + 2 * Lo; eNext.ensure(ctx, S + 1); eRank.ensure(ctx, S + 1); ePacked.ensure(ctx, S + 1); ePacked2.ensure(ctx, S + 1);
      foreach(ctx, S, EulerInitKernel{eNext.p, eRank.p});
      foreach(ctx, I, EulerLinkKernel{sortKeys.p, ordBits, itemIdx.p, objSlot.p, eNext.p, eRank.p, I});
      foreach(ctx, S, ListRankPackKernel{eNext.p, eRank.p, ePacked.p});
      const int rounds = bits_for(S);
      for (int k = 0; k < rounds; k++) {
        foreach(ctx, S, ListRankPackedKernel{ePacked.p, ePacked2.p});
        std::swap(ePacked.p, ePacked2.p); std::swap(ePacked.cap, ePacked2.cap);
      }
      foreach(ctx, S, ListRankUnpackKernel{ePacked.p, eRank.p});
      foreach(ctx, N, ListPosKernel{eRank.p, elemRow.p, objRow.p, itemIdx.p, objSlot.p, (u32)I, listPos.p});
    } else dev_memset(ctx, listPos.p, 0, (N + 1) * 4);
    checkErr(actorsNow);
    dbgMark("opset:list-ranked");
    // document order: stable LSD over (object, key rank | list position, opId within the key / element)
    perm.ensure(ctx, N + 1); pos.ensure(ctx, N + 1); sortKeys.ensure(ctx, N);
    foreach(ctx, N, IotaKernel{perm.p});
    const int fieldBits[3] = {ordBits + 1, bits_for(N), ordBits + 1};
    for (int f = 0; f < 3; f++) {
      foreach(ctx, N, DocKeyKernel{f, w, perm.p, listPos.p, keySlots.p, keySlot.p, ord, sortKeys.p});
      sortPairs(sortKeys, perm, N, fieldBits[f]);
    }
    foreach(ctx, N, InversePermKernel{perm.p, pos.p});
    dbgMark("opset:doc-ordered(enqueued)");
    // succ lists
    numPairs = numSucc + P;
    pairKey.ensure(ctx, numPairs + 1); pairSucc.ensure(ctx, numPairs + 1); pairIdx.ensure(ctx, numPairs + 1); pairPos.ensure(ctx, numPairs + 1); pairTime.ensure(ctx, numPairs + 1);
    foreach(ctx, M, DelElemCheckKernel{ops, idt, w, errWord.p});
    checkErr(actorsNow);
    foreach(ctx, numRows, OldPairsKernel{succOff.p, succ.p, pos.p, ord, pairKey.p, pairIdx.p, pairSucc.p, pairPos.p, pairTime.p});
    foreach(ctx, M, PredPairsKernel{ops, idt, pos.p, rowOfOp.p, w, elemRow.p, keySlot.p, ord, pairKey.p, pairIdx.p, pairSucc.p, pairPos.p, pairTime.p, numSucc, errWord.p});
    foreach(ctx, M, DelKeyCheckKernel{arena.p, ops, idt, w, errWord.p});
    foreach(ctx, M, IncCheckKernel{ops, idt, w, arena.p, errWord.p});
    {
      const u64 w2 = fetchErr();
      if (w2) {
        if ((w2 & 0xff) == KE_PRED_MISSING) { u64 pid = 0; d2h(ctx, &pid, o_predId.p + (w2 >> 8), 8); sync(ctx); actorIds.swap(actorsNow); std::string t = opIdText(pid); actorIds.swap(actorsNow); throw Error(AMG_ERR_RANGE, "no matching operation for pred: " + t); }
        if ((w2 & 0xff) == KE_UNKNOWN_COUNTER) { u64 oid = 0; d2h(ctx, &oid, o_id.p + (w2 >> 8), 8); sync(ctx); actorIds.swap(actorsNow); std::string t = opIdText(oid); actorIds.swap(actorsNow); throw Error(AMG_ERR_RANGE, "increment operation " + t + " for unknown counter"); }
        throwKernelError(w2, actorsNow);
      }
    }
    sortPairs(pairKey, pairIdx, numPairs, ordBits);
    foreach(ctx, numPairs, PairPosKeyKernel{pairPos.p, pairIdx.p, pairKey.p});
    sortPairs(pairKey, pairIdx, numPairs, bits_for(N));
    succCnt.ensure(ctx, N + 2); newSuccCnt.ensure(ctx, N + 2); newSuccOff.ensure(ctx, N + 2); firstNewSucc.ensure(ctx, N + 2); newSucc.ensure(ctx, numPairs + 1);
    dev_memset(ctx, succCnt.p, 0, (N + 2) * 4); dev_memset(ctx, newSuccCnt.p, 0, (N + 2) * 4); dev_memset(ctx, firstNewSucc.p, 0xff, (N + 2) * 4);
    foreach(ctx, numPairs, CountSuccKernel{pairPos.p, succCnt.p});
    foreach(ctx, numPairs, NewSuccFlagKernel{pairPos.p, pairTime.p, newSuccCnt.p});
    foreach(ctx, numPairs, FirstSuccTimeKernel{pairPos.p, pairTime.p, firstNewSucc.p});
    scan_exclusive(ctx, scanTmp, succCnt.p, newSuccOff.p, N);
    newSuccTime.ensure(ctx, numPairs + 1);
    foreach(ctx, numPairs, WriteSuccKernel{pairIdx.p, pairSucc.p, newSucc.p, pairTime.p, newSuccTime.p});
    sorted.ensure(ctx, N + 1);
    foreach(ctx, N, GatherRowsKernel{w, sorted.view(), perm.p});
    dbgMark("opset:succ+gather(enqueued)");
    timer.mark(); hostMark(); nvtx.next("patch");
    // ---------------------------------------------------------- 7. incremental patch
    if (wantPatch) {
      objPos.ensure(ctx, N + 1);
      foreach(ctx, N, ObjPosKernel{perm.p, objRow.p, pos.p, objPos.p});
      workView = w;
      buildPatch(sorted.view(), N, false, &ops, M, &idt, rowOfOp.p, pos.p, actorsNow, out, newSuccOff.p, newSucc.p);
    }
    checkErr(actorsNow);
    timer.mark(); hostMark(); nvtx.next("heads+commit");
    // heads
    dbgMark("commit:begin");
    {
      DBuf<u32>& isDep = groupLinked; isDep.ensure(ctx, G + 1); dev_memset(ctx, isDep.p, 0, (G + 1) * 4);
      foreach(ctx, B, MarkDepsKernel{applied.p, nDeps.p, depBase.p, depIdx.p, isDep.p});
      emit.ensure(ctx, B + 1); slot.ensure(ctx, B + 2); objStart.ensure(ctx, B + 1);
      foreach(ctx, B, HeadFlag2Kernel{applied.p, isDep.p, numApplied, emit.p});
      scan_exclusive(ctx, scanTmp, emit.p, slot.p, B);
      foreach(ctx, B, CompactKernel{emit.p, slot.p, objStart.p});
      // one round trip for the whole answer (HeadsPackKernel); a second one only if the call leaves more heads than the block holds
      const u32 nOld = (u32)headIdx.size(); u32 cap = 64, nh = 0; std::vector<u32> pack;
      headsPack.ensure(ctx, nOld + 1);
      if (nOld) h2d(ctx, headsPack.p, headIdx.data(), nOld * 4);
      for (;;) {
        const size_t words = 1 + (size_t)nOld + 9 * (size_t)cap;
        headsOut.ensure(ctx, words + 1); pack.resize(words);
        foreach(ctx, std::max<size_t>(std::max<size_t>(nOld, cap), 1), HeadsPackKernel{slot.p + B, objStart.p, hashes.p + numApplied * 32, appRank.p, isDep.p, headsPack.p, nOld, cap, headsOut.p});
        d2h(ctx, pack.data(), headsOut.p, words * 4); sync(ctx);
        nh = pack[0];
        if (nh <= cap) break;
        cap = nh;
      }
      std::vector<std::array<u8, 32>> hs; std::vector<u32> hi;
      for (u32 i = 0; i < nOld; i++) if (!pack[1 + i]) { hs.push_back(heads[i]); hi.push_back(headIdx[i]); }
      for (u32 k = 0; k < nh; k++) { const u32* e = pack.data() + 1 + nOld + 9 * (size_t)k; std::array<u8, 32> h; memcpy(h.data(), e, 32); hs.push_back(h); hi.push_back((u32)(numApplied + e[8])); }
      std::vector<size_t> o(hs.size()); for (size_t i = 0; i < o.size(); i++) o[i] = i;
      std::sort(o.begin(), o.end(), [&](size_t a, size_t b) { return hs[a] < hs[b]; });
      headsNow.clear(); headIdxNow.clear(); for (size_t i : o) { headsNow.push_back(hs[i]); headIdxNow.push_back(hi[i]); }
    }
  } else {
    headIdxNow = headIdx;
  }
  dbgMark("commit:heads-done");
  std::vector<std::pair<u64, UnknownRow>> unknownNow; std::set<u32> unknownIdsNow;
  if (numNew > 0 && hasUnknownColsCall) collectUnknownColumns(B, unknownNow, unknownIdsNow);   // rare: columns written by a future version (unknowncols.hpp)
  // ------------------------------------------------------------ 8. commit (nothing above mutated persistent state)
  sync(ctx);
  dbgMark("commit:synced");
  needBatch(); finishInflate();
  if (numNew > 0) {
    if (!(inOrder && numNew == B)) {   // hashes of applied changes must be contiguous in application order
      DBuf<u8>& tmp = hashTmp; tmp.ensure(ctx, numNew * 32 + 64);
      foreach(ctx, B, HashGatherKernel{hashes.p + numApplied * 32, applied.p, appRank.p, tmp.p});
      d2d(ctx, hashes.p + numApplied * 32, tmp.p, numNew * 32);
    }
    if (appliedH.empty()) {   // all applied, in order
      const u32 base0 = (u32)changes.size();
      if (!batchOriginal.empty()) for (size_t b = 0; b < B; b++) { const HostChange o = originalOf(b); if (o.len) deflatedOriginal.push_back({base0 + (u32)b, o}); }
      else { deflatedOriginal.reserve(deflatedOriginal.size() + deflIdx.size()); for (size_t k = 0; k < deflIdx.size(); k++) deflatedOriginal.push_back({base0 + deflIdx[k], inflOrig[k]}); }
      if (changes.empty()) changes.swap(batch); else changes.insert(changes.end(), batch.begin(), batch.end());
    }
    else {
      std::vector<u32> byRank(numNew);
      if (appliedH.empty()) for (size_t b = 0; b < B; b++) byRank[b] = (u32)b; else for (size_t b = 0; b < B; b++) if (appliedH[b]) byRank[appRankH[b]] = (u32)b;
      for (size_t k = 0; k < numNew; k++) {
        const u32 b = byRank[k];
        { const HostChange o = originalOf(b); if (o.len) deflatedOriginal.push_back({(u32)changes.size(), o}); }
        changes.push_back(batch[b]);
      }
    }
    dbgMark("commit:changes-recorded");
    doc.swap(sorted); numRows = N;
    std::swap(succOff.p, newSuccOff.p); std::swap(succOff.cap, newSuccOff.cap); std::swap(succ.p, newSucc.p); std::swap(succ.cap, newSucc.cap); numSucc = numPairs;
    fill32(doc.time.p, 0, N);
    for (auto& kv : unknownNow) unknownCols.byOp[kv.first] = std::move(kv.second);
    unknownCols.colIds.insert(unknownIdsNow.begin(), unknownIdsNow.end());
    loadedDoc.clear(); numApplied += numNew; actorRep = actorRepNow; actorIds = actorsNow; clock = clockNow; maxOp = maxOpNow; heads = headsNow; headIdx = headIdxNow;
    dbgMark("commit:state-swapped");
    rebuildActorTable();   // slots of actors registered in this call become permanent (first = 0)
    dbgMark("commit:actors-rebuilt");
  }
  arenaLen = cur; queue = newQueue; queueOriginal = newQueueOriginal; rb.armed = false;
  side_join(ctx); sync(ctx);
  timer.mark(); hostMark(); nvtx.next(nullptr);
  fillPatchHeader(out);
  if (isLocal && n == 1) {   // new.js:1874-1877
    std::vector<ChangeHot> m0(1); d2h(ctx, m0.data(), hot.p, sizeof(ChangeHot)); sync(ctx);
    out.hasActorSeq = true; out.actor.assign(m0[0].actorLen, '\0'); out.seq = m0[0].seq;
    if (m0[0].actorLen) { d2h(ctx, &out.actor[0], arena.p + m0[0].actorOff, m0[0].actorLen); sync(ctx); }
  }
  dbgMark("commit:end");
  lastB = B; lastM = M; lastP = P; lastBytes = cur - arenaLen0; for (auto& c : queue) lastBytes += 0 * c.len;
  finishPatch(out); dbgMark("call:patch-finished");
  timer.collect(lastPhaseMs, 12); dbgMark("call:timers-collected");
}

}  // namespace amg

namespace amg {

// Patch emission over a document table `d` in document order (N rows). wholeDoc = getPatch semantics
// (new.js:1604-1635), otherwise incremental semantics for the batch `ops` (new.js:884-1040, 1461-1528).
// Uses succCnt (per position) and, in incremental mode, newSuccCnt / firstNewSucc / objPos.
inline void Engine::buildPatch(DocRows d, size_t N, bool wholeDoc, const OpRows* ops, size_t numOps, const IdTable* idt, const u32* rowOfOpD, const u32* posD,
                               const std::vector<std::string>& actorsNow, PatchOut& out, const u32* succOffD, const u64* succD) {
  out.numProps = out.numEdits = 0; out.bigEnd = 0;
  if (N == 0) return;
  // groups (map key / list element) and their visibility
  head.ensure(ctx, N + 1); headScan.ensure(ctx, N + 2); groupOf.ensure(ctx, N + 1);
  foreach(ctx, N, GroupHeadKernel{arena.p, d, head.p});
  scan_exclusive(ctx, scanTmp, head.p, headScan.p, N);
  isObjHead.ensure(ctx, N + 1); objIdx.ensure(ctx, N + 2);
  foreach(ctx, N, ObjHeadKernel{d, isObjHead.p});
  scan_exclusive(ctx, scanTmp, isObjHead.p, objIdx.p, N);
  u32 numGroups32 = 0, numObjs32 = 0; readU32x2(headScan.p + N, objIdx.p + N, &numGroups32, &numObjs32);
  const size_t numGroups = numGroups32, numObjs = numObjs32;
  groupRows.ensure(ctx, numGroups + 1); groupVisible.ensure(ctx, numGroups + 1); groupFirst.ensure(ctx, numGroups + 1); groupTouched.ensure(ctx, numGroups + 1);
  DBuf<u32>& groupLinkedB = linkDone;   // linkDone doubles as per-group linked flags storage below (separate buffers)
  (void)groupLinkedB;
  dev_memset(ctx, groupRows.p, 0, (numGroups + 1) * 4); dev_memset(ctx, groupVisible.p, 0, (numGroups + 1) * 4); dev_memset(ctx, groupTouched.p, 0, (numGroups + 1) * 4);
  groupHasChild.ensure(ctx, numGroups + 1); dev_memset(ctx, groupHasChild.p, 0, (numGroups + 1) * 4);
  foreach(ctx, N, GroupStatsKernel{headScan.p, head.p, succCnt.p, d, groupOf.p, groupRows.p, groupVisible.p, groupFirst.p, errWord.p, 0, groupHasChild.p});
  // objects in document order
  objStart.ensure(ctx, numObjs + 2);
  foreach(ctx, N, ObjStartKernel{isObjHead.p, objIdx.p, objStart.p, N});
  { const u32 nn = (u32)N; h2d(ctx, objStart.p + numObjs, &nn, 4); }
  emit.ensure(ctx, N + 1); marker.ensure(ctx, N + 1); slot.ensure(ctx, N + 2);
  groupLinked.ensure(ctx, std::max(numGroups, numApplied + 1) + 2);
  dev_memset(ctx, groupLinked.p, 0, (numGroups + 1) * 4);
  Ord ordNow{actorRank.p, bits_for(actorsNow.size() > 1 ? actorsNow.size() - 1 : 1)};
  ListCtx lctx{d, succCnt.p, newSuccCnt.p, firstNewSucc.p, groupOf.p, groupFirst.p, groupRows.p, arena.p, succOffD, succD, newSuccTime.p};
  MapGroupCtx mg{arena.p, ops ? *ops : OpRows{}, opAt.p, numOps, pass.p};
  bool anyListLink = false;
  auto listGroups = [&](int pass) {
    return ListGroupKernel{pass, mg, opGroupHead.p, *idt, rowOfOpD, posD, lctx, gCount.p, gElem.p, gT1.p, gQOrd.p, nQ.p, elemHasRecs.p, elemMinT.p,
                           itemBase.p, objIdx.p, objStart.p, items.p, domTw.p, domW.p, oldVisScan.p, runHeadFlag.p, runScan.p, runStart.p, gBase.p, qIndex.p, editOut.p, editElem.p, editObjKey.p, editElemPos.p, editRowPos.p, errWord.p};
  };
  if (!wholeDoc) {
    // op groups of the batch (new.js:1085-1138), then what each list group nets out to
    nQ.ensure(ctx, N + 1); elemHasRecs.ensure(ctx, N + 1); listLinkTime.ensure(ctx, N + 1); elemMinT.ensure(ctx, N + 1);
    dev_memset(ctx, elemMinT.p, 0xff, (N + 1) * 4);
    dev_memset(ctx, nQ.p, 0, (N + 1) * 4); dev_memset(ctx, elemHasRecs.p, 0, (N + 1) * 4); dev_memset(ctx, listLinkTime.p, 0xff, (N + 1) * 4);
    if (numOps > 0) {
      opAt.ensure(ctx, numOps + 1); runHead.ensure(ctx, numOps + 1); opGroupHead.ensure(ctx, numOps + 1);
      for (DBuf<u32>* b : {&gCount, &gElem, &gT1, &gQOrd, &gBase, &qIndex}) b->ensure(ctx, numOps + 2);
      mg.opAt = opAt.p;
      foreach(ctx, numOps, OpAtTimeKernel{ops->time, opAt.p});
      foreach(ctx, numOps, RunHeadKernel{mg, runHead.p});
      foreach(ctx, numOps, GroupSplitKernel{mg, runHead.p, opGroupHead.p});
      foreach(ctx, numOps, listGroups(0));
      runHeadFlag.ensure(ctx, numOps + 1); runScan.ensure(ctx, numOps + 2); runStart.ensure(ctx, numOps + 2); elemFollower.ensure(ctx, N + 1);
      dev_memset(ctx, elemFollower.p, 0, (N + 1) * 4);
      foreach(ctx, numOps, FollowerFlagKernel{mg, opGroupHead.p, lctx, gElem.p, gT1.p, gCount.p, nQ.p, runHeadFlag.p, elemFollower.p});
      scan_exclusive(ctx, scanTmp, runHeadFlag.p, runScan.p, numOps);
      foreach(ctx, numOps, RunStartKernel{runHeadFlag.p, runScan.p, runStart.p});
      foreach(ctx, 1, RunEndKernel{runScan.p, runStart.p, (u32)numOps});
    }
    objTouchedAt.ensure(ctx, N + 1); linkDone.ensure(ctx, N + 1);
    dev_memset(ctx, objTouchedAt.p, 0xff, (N + 1) * 4); dev_memset(ctx, linkDone.p, 0xff, (N + 1) * 4); dev_memset(ctx, flagWord.p, 0, 16);
    foreach(ctx, N, TouchKernel{d, groupOf.p, firstNewSucc.p, groupTouched.p, objTouchedAt.p, objPos.p, flagWord.p + 2});
    u32 linkChanged = 1, anyLink32 = 0;
    for (int iter = 0; iter < 1000 && linkChanged; iter++) {   // three sweeps per host round trip (object nesting is shallow)
      LinkKernel lk{d, groupOf.p, groupHasChild.p, groupFirst.p, objPos.p, groupLinked.p, objTouchedAt.p, flagWord.p + 2, linkDone.p, flagWord.p, listLinkTime.p, flagWord.p + 3};
      foreach(ctx, N, lk); foreach(ctx, N, lk);
      dev_memset(ctx, flagWord.p, 0, 4);
      foreach(ctx, N, lk);
      readU32x2(flagWord.p, flagWord.p + 3, &linkChanged, &anyLink32);
    }
    anyListLink = anyLink32 != 0;
  }
  // ---- map props
  DBuf<u32>& groupEmitted = elemVis;   // scratch reuse (list edits re-initialise it later)
  groupEmitted.ensure(ctx, std::max(numGroups, N) + 2); dev_memset(ctx, groupEmitted.p, 0, (numGroups + 1) * 4);
  finalTime.ensure(ctx, numGroups + 1); gBound.ensure(ctx, numGroups + 1); gFailed.ensure(ctx, numGroups + 1); memberFinal.ensure(ctx, N + 1);
  dev_memset(ctx, finalTime.p, 0, (numGroups + 1) * 4);
  if (!wholeDoc && numOps > 0) {
    dev_memset(ctx, memberFinal.p, 0, (N + 1) * 4); dev_memset(ctx, gFailed.p, 0, (numGroups + 1) * 4);
    for (int pass = 0; pass < 2; pass++)
      foreach(ctx, numOps, GroupFinalKernel{pass, mg, opGroupHead.p, *idt, rowOfOpD, posD, groupOf.p, workView, ordNow, finalTime.p, gBound.p, gFailed.p, memberFinal.p});
  }
  counterLast.ensure(ctx, N + 1); counterTotal.ensure(ctx, N + 1); counterOwner.ensure(ctx, N + 1); dev_memset(ctx, counterOwner.p, 0xff, (N + 1) * 4);
  foreach(ctx, N, CounterKernel{arena.p, d, succOffD, succD, groupOf.p, groupFirst.p, groupRows.p, counterLast.p, counterTotal.p, counterOwner.p});
  foreach(ctx, N, PropFlagKernel{d, groupOf.p, groupTouched.p, groupLinked.p, succCnt.p, wholeDoc ? 1 : 0, finalTime.p, gBound.p, gFailed.p, memberFinal.p, ordNow, emit.p, groupEmitted.p, counterLast.p});
  foreach(ctx, N, PropMarkerKernel{d, groupOf.p, groupTouched.p, head.p, groupEmitted.p, wholeDoc ? 1 : 0, emit.p, marker.p});
  scan_exclusive(ctx, scanTmp, emit.p, slot.p, N);
  const size_t numProps = readU32(slot.p + N);
  propOut.ensure(ctx, numProps + 1);
  foreach(ctx, N, PropEmitKernel{d, emit.p, marker.p, slot.p, propOut.p, counterLast.p, counterTotal.p});
  if (curTimer) { curTimer->mark(); curHostMark(); }
  NvtxPhases nvtxPatch; nvtxPatch.next("patch:list-edits");
  // ---- list edits
  size_t numEdits = 0; bool shipElem = true;
  if (wholeDoc) {
    elemVis.ensure(ctx, N + 1); elemVisScan.ensure(ctx, N + 2); rowEmit.ensure(ctx, N + 1); firstVis.ensure(ctx, numGroups + 1);
    rowClass.ensure(ctx, N + 1); firstBare.ensure(ctx, numGroups + 1);
    dev_memset(ctx, firstVis.p, 0xff, (numGroups + 1) * 4); dev_memset(ctx, firstBare.p, 0xff, (numGroups + 1) * 4);
    foreach(ctx, N, ListRowClassKernel{d, groupOf.p, succCnt.p, counterOwner.p, rowClass.p, firstVis.p, firstBare.p});
    foreach(ctx, N, ListVisFlagKernel{d, groupOf.p, groupVisible.p, head.p, succCnt.p, elemVis.p, rowEmit.p, rowClass.p, firstVis.p, firstBare.p});
    scan_exclusive(ctx, scanTmp, elemVis.p, elemVisScan.p, N);
    scan_exclusive(ctx, scanTmp, rowEmit.p, slot.p, N);
    numEdits = readU32(slot.p + N);
    editOut.ensure(ctx, numEdits + 1); editElem.ensure(ctx, numEdits + 1);
    foreach(ctx, N, DocEditEmitKernel{d, rowEmit.p, slot.p, elemVisScan.p, objIdx.p, objStart.p, groupOf.p, groupFirst.p, rowClass.p, firstVis.p, firstBare.p, editOut.p, counterOwner.p, counterTotal.p});
    foreach(ctx, N, EditElemKernel{d, rowEmit.p, slot.p, groupOf.p, groupFirst.p, editElem.p});
  } else {
    size_t numGroupRecs = 0, numLink = 0, numLive = 0;
    if (numOps > 0) { scan_exclusive(ctx, scanTmp, gCount.p, gBase.p, numOps); numGroupRecs = readU32(gBase.p + numOps); }
    DBuf<u32>& elemHasLive = elemHasRecs; dev_memset(ctx, elemHasLive.p, 0, (N + 1) * 4);
    auto ensureEdits = [&](size_t n) {
      editOut.ensure(ctx, n + 1, numLive); editElem.ensure(ctx, n + 1, numLive); editObjKey.ensure(ctx, n + 1, numLive); editElemPos.ensure(ctx, n + 1);
      editOut2.ensure(ctx, n + 1); editElem2.ensure(ctx, n + 1); editElemPos2.ensure(ctx, n + 1); editObjKey2.ensure(ctx, n + 1); editRowPos.ensure(ctx, n + 1); editRowPos2.ensure(ctx, n + 1);
      sortKeys.ensure(ctx, n + 1); sortVals.ensure(ctx, n + 1);
    };
    // ---- A. records of the op groups: indexes, emission, per-object order, pops / coalescing, compaction
    if (numGroupRecs > 0) {
      // list index of a group = elements in front that were visible before the batch (one prefix sum)
      //                       + net visibility changes in front that earlier groups of the batch made (dominance count)
      nItems.ensure(ctx, N + 1); itemBase.ensure(ctx, N + 2); oldVisScan.ensure(ctx, N + 2);
      foreach(ctx, N, OldVisFlagKernel{lctx, head.p, nItems.p});
      scan_exclusive(ctx, scanTmp, nItems.p, oldVisScan.p, N);
      foreach(ctx, N, DomItemCountKernel{nQ.p, elemFollower.p, nItems.p});
      scan_exclusive(ctx, scanTmp, nItems.p, itemBase.p, N);
      const size_t T = readU32(itemBase.p + N);
      items.ensure(ctx, T + 1); items2.ensure(ctx, T + 1); zwScan.ensure(ctx, T + 2); domTw.ensure(ctx, T + 1); domTw2.ensure(ctx, T + 1); domW.ensure(ctx, T + 1); domW2.ensure(ctx, T + 1);
      foreach(ctx, numOps, listGroups(1));
      const int tbits = bits_for(numOps + 1);
#ifdef AMG_EMU
      const int localBits = 0;
#else
      const int localBits = std::min(tbits, DOM_LOCAL_BITS);   // the last levels run inside shared memory (k_dom_local)
#endif
      for (int bit = tbits - 1; bit >= localBits; bit--) {
        scan_exclusive64(ctx, scanTmp, DomScanInput{domTw.p, domW.p, bit}, zwScan.p, T);
        foreach(ctx, T, DomLevelKernel{items.p, items2.p, domTw2.p, domW2.p, zwScan.p, bit});
        std::swap(items.p, items2.p); std::swap(items.cap, items2.cap); std::swap(domTw.p, domTw2.p); std::swap(domTw.cap, domTw2.cap); std::swap(domW.p, domW2.p); std::swap(domW.cap, domW2.cap);
      }
#ifdef AMG_EMU
      foreach(ctx, T, DomResultKernel{items.p, qIndex.p});
#else
      {   // partition heads -> compact list (device-side count), then one CTA per partition
        DBuf<u32>& partFlag = domTw2; DBuf<u32>& partScan = nItems; DBuf<u32>& partHead = itemBase;   // scratch that is free by now
        partScan.ensure(ctx, T + 2); partHead.ensure(ctx, T + 2);
        foreach(ctx, T, DomPartHeadKernel{items.p, partFlag.p});
        scan_exclusive(ctx, scanTmp, partFlag.p, partScan.p, T);
        foreach(ctx, T, CompactKernel{partFlag.p, partScan.p, partHead.p});
        if (!domLocalReady) { CUDA_CHECK(cudaFuncSetAttribute(k_dom_local, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DomLocalSmem))); domLocalReady = true; }
        k_dom_local<<<ctx.numSMs * 2, 256, sizeof(DomLocalSmem), ctx.stream>>>(items.p, partHead.p, partScan.p + T, qIndex.p, localBits, errWord.p);
        CUDA_CHECK(cudaGetLastError()); ctx.launches++;
      }
#endif
      if (curTimer) { curTimer->mark(); curHostMark(); }
      ensureEdits(numGroupRecs);
      foreach(ctx, numOps, listGroups(2));
      // the records are in application order by construction: one stable sort by object gives the per-object edit lists
      foreach(ctx, numGroupRecs, EditKeyKernel{editObjKey.p, editOut.p, sortKeys.p, sortVals.p});
      sortPairs(sortKeys, sortVals, numGroupRecs, bits_for(numObjs));
      foreach(ctx, numGroupRecs, EditGatherKernel{editOut.p, editElem.p, editElemPos.p, editObjKey.p, sortVals.p, editOut2.p, editElem2.p, editElemPos2.p, editObjKey2.p});
      foreach(ctx, numGroupRecs, GatherU32Kernel{editRowPos.p, sortVals.p, editRowPos2.p});
      for (DBuf<u32>* b : {&editKind, &editPred, &editDead, &editMerge, &editMulti, &editLive}) b->ensure(ctx, numGroupRecs + 2);
      dev_memset(ctx, editDead.p, 0, (numGroupRecs + 1) * 4); dev_memset(ctx, editMulti.p, 0, (numGroupRecs + 1) * 4);
      foreach(ctx, numGroupRecs, EditFixKernel{editOut2.p, editElemPos2.p, editKind.p, editPred.p, editDead.p, numGroupRecs});
      foreach(ctx, numGroupRecs, EditMergeKernel{editOut2.p, editElem2.p, editKind.p, editPred.p, editMerge.p, editMulti.p});
      dev_memset(ctx, flagWord.p, 0, 4);
      foreach(ctx, numGroupRecs, EditLiveKernel{editDead.p, editLive.p, editOut2.p, editElem2.p, editKind.p, flagWord.p, editElemPos2.p, elemHasLive.p, editRowPos2.p, succCnt.p, counterLast.p});
      DBuf<u32>& liveSlot = editPred;   // pred is consumed by now
      scan_exclusive(ctx, scanTmp, editLive.p, liveSlot.p, numGroupRecs);
      u32 numLive32 = 0, needElem32 = 0; readU32x2(liveSlot.p + numGroupRecs, flagWord.p, &numLive32, &needElem32);
      numLive = numLive32; shipElem = needElem32 != 0;
      foreach(ctx, numGroupRecs, EditCompactKernel{editOut2.p, editElem2.p, editDead.p, liveSlot.p, editKind.p, editMerge.p, editMulti.p, editOut.p, editElem.p, editObjKey2.p, editObjKey.p});
    } else if (curTimer) { curTimer->mark(); curHostMark(); }
    // ---- B. setupPatches link edits on list parents: only for elements that did not keep an edit of their own; appended
    //         behind the object's other edits in the order the child objects were first touched
    if (anyListLink) {
      DBuf<u32>& linkCount = rowEmit; DBuf<u32>& linkBase = slot;
      elemVis.ensure(ctx, N + 1); elemVisScan.ensure(ctx, N + 2); linkCount.ensure(ctx, N + 1); linkBase.ensure(ctx, N + 2);
      foreach(ctx, N, ListVisFlagKernel{d, groupOf.p, groupVisible.p, head.p, succCnt.p, elemVis.p, linkCount.p, nullptr, nullptr, nullptr});
      scan_exclusive(ctx, scanTmp, elemVis.p, elemVisScan.p, N);   // index of a linked element = visible elements before it once the whole batch is applied
      foreach(ctx, N, ListLinkKernel{0, lctx, listLinkTime.p, elemVisScan.p, objIdx.p, objStart.p, linkCount.p, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, elemHasLive.p});
      scan_exclusive(ctx, scanTmp, linkCount.p, linkBase.p, N);
      numLink = readU32(linkBase.p + N);
      if (numLink > 0) {
        const size_t total = numLive + numLink;
        ensureEdits(total); editTime.ensure(ctx, total + 1);
        foreach(ctx, N, ListLinkKernel{1, lctx, listLinkTime.p, elemVisScan.p, objIdx.p, objStart.p, linkCount.p, linkBase.p, (u32)numLive, editOut.p, editElem.p, editObjKey.p, editElemPos.p, editTime.p, elemHasLive.p});
        // order: surviving group records as they are, then the link edits by first-touch time; then stably by object
        foreach(ctx, numLive, OffsetIotaKernel{sortVals.p, 0});
        if (numLink > 1) {
          DBuf<u64>& k2 = pairKey; DBuf<u32>& v2 = pairIdx; k2.ensure(ctx, numLink + 1); v2.ensure(ctx, numLink + 1);
          foreach(ctx, numLink, EditTimeKeyAtKernel{editTime.p, (u32)numLive, k2.p, v2.p});
          sortPairs(k2, v2, numLink, 32);
          d2d(ctx, sortVals.p + numLive, v2.p, numLink * 4);
        } else foreach(ctx, numLink, OffsetIotaKernel{sortVals.p + numLive, (u32)numLive});
        foreach(ctx, total, GatherToU64Kernel{editObjKey.p, sortVals.p, sortKeys.p});
        sortPairs(sortKeys, sortVals, total, bits_for(numObjs));
        foreach(ctx, total, EditGatherKernel{editOut.p, editElem.p, nullptr, nullptr, sortVals.p, editOut2.p, editElem2.p, nullptr, nullptr});
        std::swap(editOut.p, editOut2.p); std::swap(editOut.cap, editOut2.cap); std::swap(editElem.p, editElem2.p); std::swap(editElem.cap, editElem2.cap);
      }
    }
    numEdits = numLive + numLink;
  }
  if (wholeDoc && numEdits > 0) foreach(ctx, numEdits, RunFlagKernel{editOut.p, editElem.p, numEdits});
  out.numProps = numProps; out.numEdits = numEdits;
  out.propsOff = PATCH_HDR_WORDS * 8; out.editsOff = out.propsOff + numProps * sizeof(PropRec);
  size_t end;
  if (shipElem) { out.elemOff = out.editsOff + numEdits * sizeof(EditRec); end = out.elemOff + numEdits * 8; }
  else { out.elemOff = 0; end = out.editsOff + numEdits * sizeof(EditRec); }   // elemOff 0: every insert's elemId is its opId
  // key and value bytes of the records: gathered behind them, offsets rewritten to positions inside the patch
  const size_t nRec = numProps + numEdits; size_t nBytes = 0;
  out.valBytesOff = end;
  if (nRec > 0) {
    patchByteLen.ensure(ctx, nRec + 1); patchByteOff.ensure(ctx, nRec + 2);
    foreach(ctx, nRec, PatchBytesCountKernel{propOut.p, numProps, editOut.p, patchByteLen.p});
    scan_exclusive(ctx, scanTmp, patchByteLen.p, patchByteOff.p, nRec);
    nBytes = readU32(patchByteOff.p + nRec);
    if ((u64)end + nBytes >= 0xfff00000ULL) throw Error(AMG_ERR_UNSUPPORTED, "amgpu: patch larger than 4 GiB");
    patchBytesD.ensure(ctx, nBytes + 8);
    foreach(ctx, nRec, PatchBytesGatherKernel{arena.p, propOut.p, numProps, editOut.p, patchByteOff.p, (u32)out.valBytesOff, patchBytesD.p, errWord.p});
    {   // a value that the reference's decodeValue refuses (it decodes every value that reaches a patch, columnar.js:300-329)
      const u64 ew = fetchErr();
      if ((ew & 0xff) == KE_FLOAT_LEN) { u32 l = 0; d2h(ctx, &l, patchByteLen.p + (ew >> 8), 4); sync(ctx); const size_t i = (size_t)(ew >> 8); u32 keyLen = 0; if (i < numProps) { PropRec r; d2h(ctx, &r, propOut.p + i, sizeof(PropRec)); sync(ctx); keyLen = r.keyLen == 0xffffffffu ? 0 : r.keyLen; } throw Error(AMG_ERR_RANGE, "Invalid length for floating point number: " + std::to_string(l - keyLen)); }
      if (ew) throwKernelError(ew, actorsNow);
    }
  }
  out.valBytesLen = nBytes; out.bigEnd = (out.valBytesOff + nBytes + 7) & ~(size_t)7;
  patchBuf.ensure(out.bigEnd + 4096);
  // the copy-out runs on the side stream: the caller joins it before reading the patch, later kernels overlap it
  side_fork(ctx);
  d2h_side(ctx, patchBuf.p + out.propsOff, propOut.p, numProps * sizeof(PropRec));
  if (numEdits > 0) { d2h_side(ctx, patchBuf.p + out.editsOff, editOut.p, numEdits * sizeof(EditRec)); if (shipElem) d2h_side(ctx, patchBuf.p + out.elemOff, editElem.p, numEdits * 8); }
  if (nBytes > 0) d2h_side(ctx, patchBuf.p + out.valBytesOff, patchBytesD.p, nBytes);
}

inline void Engine::getPatch(PatchOut& out) {
  dev_memset(ctx, errWord.p, 0, 16); errSnapLaunches = ~0ull;
  succCnt.ensure(ctx, numRows + 2);
  foreach(ctx, numRows, SuccCntFromOffKernel{succOff.p, succCnt.p});
  struct SideJoin { Ctx& c; ~SideJoin() { side_join(c); } } sideJoin{ctx};
  buildPatch(doc.view(), numRows, true, nullptr, 0, nullptr, nullptr, nullptr, actorIds, out, succOff.p, succ.p);
  side_join(ctx); sync(ctx);
  checkErr(actorIds);
  fillPatchHeader(out);
  finishPatch(out);
}

}  // namespace amg

namespace amg {

inline RawRows Engine::rawRows() {
  return RawRows{r_objActor.p, r_objCtr.p, r_keyActor.p, r_keyCtr.p, r_keyStrOff.p, r_keyStrLen.p, r_insert.p, r_action.p, r_valLen.p, r_valOff.p, r_predNum.p, r_predOff.p, r_predActor.p, r_predCtr.p};
}
// Fused header parse + column expansion of B changes (chOff / chLen are on the device). The raw row tables are sized from
// what earlier calls needed (else from the batch size); the kernel never writes outside them and reports an overflow.
inline DecodeTilesArgs Engine::decodeArgs(const u8* arenaP, size_t B, size_t batchBytes) {
  hot.ensure(ctx, B + 1); nOps.ensure(ctx, B + 1); nPreds.ensure(ctx, B + 1); nDeps.ensure(ctx, B + 1); nActors.ensure(ctx, B + 1);
  rawBase.ensure(ctx, B + 2); rawPredBase.ensure(ctx, B + 2); decErr.ensure(ctx, B + 1);
  decCursor.ensure(ctx, 8); decDirect.ensure(ctx, B + 2);   // decCursor: one 64-byte block = cursor (2 x u64), totals (4 x u32), direct count, done count
  const size_t wantRows = std::max(decWantRows, B + B / 4 + batchBytes / 256 + 1024), wantPreds = std::max(decWantPreds, B + B / 4 + batchBytes / 256 + 1024);
  for (DBuf<u32>* b : {&r_objActor, &r_objCtr, &r_keyActor, &r_keyCtr, &r_keyStrOff, &r_keyStrLen, &r_insert, &r_action, &r_valLen, &r_valOff, &r_predNum, &r_predOff}) b->ensure(ctx, wantRows + 1);
  r_predActor.ensure(ctx, wantPreds + 1); r_predCtr.ensure(ctx, wantPreds + 1);
  decRowCap = wantRows; decPredCap = wantPreds;
  return DecodeTilesArgs{arenaP, chOff.p, chLen.p, (u32)B, 0u, hot.p, nOps.p, nPreds.p, nDeps.p, nActors.p, rawBase.p, rawPredBase.p, decErr.p, rawRows(), (u32)wantRows, (u32)wantPreds,
                         (unsigned long long*)decCursor.p, decTotalsPtr(), errWord.p, 0u, decDirect.p, decTotalsPtr() + 4};
}
// the whole batch in one go (bytes resident): the range launch, the list launch over the inflated changes (numDefl of them,
// bytes from arena offset deflStart on), then the changes outside their tile's window
inline void Engine::runDecodeTiles(const u8* arenaP, size_t B, size_t batchBytes, const u32* deflListP, size_t numDefl, size_t deflStart) {
  DecodeTilesArgs a = decodeArgs(arenaP, B, batchBytes);
  if (numDefl > 0) a.skipFrom = (u32)deflStart;
  decode_tiles_begin(ctx, a); decode_tiles_range(ctx, a, 0, (u32)B);
  if (numDefl > 0) decode_tiles_list(ctx, a, deflListP, (u32)numDefl);
  decode_tiles_finish(ctx, a, B);
}
inline bool Engine::decodeOverflowed(const u32 totals[4]) {
  if (!totals[2]) return false;
  if (totals[0] >= (1u << 29)) throw Error(AMG_ERR_UNSUPPORTED, "amgpu: more than 2^29 operations in one call");
  if (totals[1] >= (1u << 30)) throw Error(AMG_ERR_UNSUPPORTED, "amgpu: more than 2^30 predecessors in one call");
  decWantRows = (size_t)totals[0] + 1024; decWantPreds = (size_t)totals[1] + 1024;
  return true;
}

// Values of columns with unknown ids in the changes this call applies (reference new.js:1406-1424 keeps them in the document).
// Host work on a rare path: headers, op counts and actor maps come back from the device, the change bytes from the arena.
inline void Engine::collectUnknownColumns(size_t B, std::vector<std::pair<u64, UnknownRow>>& out, std::set<u32>& ids) {
  std::vector<ChangeHot> hh(B); std::vector<u32> nops(B), amb(B + 1), nact(B); std::vector<u8> ap(B);
  d2h(ctx, hh.data(), hot.p, B * sizeof(ChangeHot)); d2h(ctx, nops.data(), nOps.p, B * 4); d2h(ctx, amb.data(), amapBase.p, (B + 1) * 4); d2h(ctx, nact.data(), nActors.p, B * 4); d2h(ctx, ap.data(), applied.p, B); sync(ctx);
  std::vector<u32> am(amb[B] + 1); if (amb[B]) { d2h(ctx, am.data(), amap.p, (size_t)amb[B] * 4); sync(ctx); }
  for (size_t b = 0; b < B; b++) {
    if (!ap[b] || nops[b] == 0 || hh[b].dataOff <= hh[b].dirOff) continue;
    const ChangeHot& h = hh[b];
    std::vector<u8> bytes(h.len); d2h(ctx, bytes.data(), arena.p + h.off, h.len); sync(ctx);
    std::vector<std::array<u32, 3>> cols; bool any = false;
    { ByteReader d(bytes.data(), h.dirOff - h.off, h.dataOff - h.off); u32 pos = h.dataOff - h.off;
      while (!d.done() && !d.err) { const u32 id = (u32)d.uleb(), l = (u32)d.uleb(); cols.push_back({id, pos, l}); if (!is_known_change_column(id)) any = true; pos += l; } }
    if (!any) continue;
    const u32 author = am[amb[b]];
    const u32 e = read_unknown_columns(bytes.data(), cols, nops[b], is_known_change_column, [&](size_t i, UnknownRow& row) {
      for (auto& kv : row) {
        ids.insert(kv.first);
        if ((kv.first & 7) == 1) for (auto& v : kv.second) if (!v.isNull) {   // ACTOR_ID: change-local index -> document actor index (new.js:586-588)
          if ((u64)v.num >= nact[b]) throw Error(AMG_ERR_RANGE, "actor index out of range");
          v.num = am[amb[b] + (u32)v.num];
        }
      }
      out.emplace_back(pack_id(h.startOp + i, author), row);
    });
    if (e == KE_UNSUPPORTED_OP) throw Error(AMG_ERR_RANGE, "unexpected VALUE_RAW column");
    if (e) throwKernelError(((u64)b << 8) | e, actorIds);
  }
}

// save(): document columns of the unknown ids, rows in document order; a row without a value in a column contributes null
// (a group cardinality of 0 for grouped columns), as the reference's decoders yield for columns a block never had
// (new.js:1418-1420 makeDecoders over the widened column list).
inline void Engine::appendUnknownDocColumns(std::vector<std::pair<u32, std::string>>& cols) {
  const size_t N = numRows; std::vector<u64> ids(N); d2h(ctx, ids.data(), doc.id.p, N * 8); sync(ctx);
  for (u32 id : unknownCols.colIds) {
    std::vector<UnknownValue> vals; vals.reserve(N);
    const bool grouped = (id & 7) != 0 && (unknownCols.colIds.count((id & ~15u)) != 0 || (id >> 4) == 7 || (id >> 4) == 8);   // a member of a group with a GROUP_CARD column: a row without values has cardinality 0
    for (size_t i = 0; i < N; i++) {
      auto it = unknownCols.byOp.find(ids[i]);
      const std::vector<UnknownValue>* v = nullptr;
      if (it != unknownCols.byOp.end()) { auto c = it->second.find(id); if (c != it->second.end()) v = &c->second; }
      if (v) vals.insert(vals.end(), v->begin(), v->end());
      else if ((id & 7) == 0) { UnknownValue z; z.isNull = false; z.num = 0; vals.push_back(z); }   // GROUP_CARD: `readValue() || 0` (new.js:581)
      else if (!grouped && (id & 7) != 7) vals.push_back(UnknownValue());   // null (raw bytes: nothing)
    }
    cols.emplace_back(id, encode_unknown_column(id, vals));
  }
}

// Columns of bulk changes (>= HUGE_CHANGE_OPS ops) through the parallel column decoders. largeList holds the large changes of
// the batch (at most 8 here). hugeDone[k * NCOLS + col] = 1 tells DecodeColumnKernel that column `col` of large change k is done.
inline u32 Engine::decodeHugeChanges(const RawRows& raw, size_t numLarge) {
  static const u32 HUGE_CHANGE_OPS = 4096;
  hugeDone.ensure(ctx, numLarge * NCOLS + 1); dev_memset(ctx, hugeDone.p, 0, (numLarge * NCOLS + 1) * 4);
  std::vector<u32> list(numLarge); d2h(ctx, list.data(), largeList.p, numLarge * 4); sync(ctx);
  u32 any = 0;
  for (size_t k = 0; k < numLarge; k++) {
    const u32 c = list[k]; ChangeHot h; u32 n = 0, np = 0, rb = 0, rpb = 0;
    d2h(ctx, &h, hot.p + c, sizeof(ChangeHot)); d2h(ctx, &n, nOps.p + c, 4); d2h(ctx, &np, nPreds.p + c, 4); d2h(ctx, &rb, rawBase.p + c, 4); d2h(ctx, &rpb, rawPredBase.p + c, 4); sync(ctx);
    if (n < HUGE_CHANGE_OPS || h.dataOff <= h.dirOff || h.dataOff - h.dirOff > 4096) continue;
    std::vector<u8> dir(h.dataOff - h.dirOff); d2h(ctx, dir.data(), arena.p + h.dirOff, dir.size()); sync(ctx);
    u32 cOff[NCOLS] = {0}, cLen[NCOLS] = {0}; bool have[NCOLS] = {false};
    { ByteReader d(dir.data(), 0, (u32)dir.size()); u32 pos = h.dataOff;
      while (!d.done() && !d.err) { const u32 id = (u32)d.uleb(), l = (u32)d.uleb(); const int ix = col_index_of(id); if (ix >= 0) { cOff[ix] = pos; cLen[ix] = l; have[ix] = true; } pos += l; } }
    std::vector<u32> done(NCOLS, 0);
    auto bytesOf = [&](int ix) { return arena.p + cOff[ix]; };
    struct Plan { int col; u32* out; size_t cnt; };
    const Plan plan[] = {{CX_OBJ_ACTOR, raw.objActor + rb, n}, {CX_OBJ_CTR, raw.objCtr + rb, n}, {CX_KEY_ACTOR, raw.keyActor + rb, n}, {CX_KEY_CTR, raw.keyCtr + rb, n},
                         {CX_INSERT, raw.insert + rb, n}, {CX_ACTION, raw.action + rb, n}, {CX_VAL_LEN, raw.valLen + rb, n}, {CX_PRED_NUM, raw.predNum + rb, n},
                         {CX_PRED_ACTOR, raw.predActor + rpb, np}, {CX_PRED_CTR, raw.predCtr + rpb, np}};
    for (const Plan& pl : plan) {
      if (!have[pl.col] || cLen[pl.col] == 0 || pl.cnt == 0) continue;   // absent / empty: the serial path fills the defaults
      const u8* bytes = bytesOf(pl.col); const u32 len = cLen[pl.col]; bool ok = false;
      switch (pl.col) {
        case CX_OBJ_ACTOR: case CX_OBJ_CTR: case CX_KEY_ACTOR: case CX_ACTION: case CX_PRED_ACTOR: ok = parCols.toU32(bytes, len, pl.cnt, pl.out); break;
        case CX_KEY_CTR: case CX_PRED_CTR: ok = parCols.deltaToU32(bytes, len, pl.cnt, pl.out); break;
        case CX_INSERT: ok = parCols.boolean(bytes, len, pl.cnt, pl.out); break;
        case CX_VAL_LEN: { u64 sum = 0; ok = parCols.lenColumn(bytes, len, pl.cnt, raw.valLen + rb, raw.valOff + rb, have[CX_VAL_RAW] ? cOff[CX_VAL_RAW] : 0, &sum) && sum <= (have[CX_VAL_RAW] ? cLen[CX_VAL_RAW] : 0); } break;
        case CX_PRED_NUM: { u64 sum = 0; ok = parCols.countColumn(bytes, len, pl.cnt, raw.predNum + rb, raw.predOff + rb, &sum) && sum == np; if (ok && rpb) foreach(ctx, pl.cnt, PcAddBaseKernel{raw.predOff + rb, rpb}); } break;
        default: break;
      }
      if (ok) { done[pl.col] = 1; any |= 1u << pl.col; }
    }
    h2d(ctx, hugeDone.p + k * NCOLS, done.data(), NCOLS * 4); sync(ctx);
    if (getenv("AMG_PAR_DOC_TRACE")) fprintf(stderr, "amgpu decode: bulk change %u (%u ops, %u preds): columns expanded in parallel: mask %04x\n", c, n, np, any);
  }
  return any;
}

// Re-runs the decode kernels over the last applied batch (bytes resident in HBM) and times them with CUDA events.
// msParse = the fused decode (k_decode_tiles: header parse + expansion of every change of up to SMALL_CHANGE_OPS ops),
// msDec = DecodeColumnKernel over the larger changes (0 when the batch has none).
inline void Engine::benchDecode(int iters, float* msSha, float* msParse, float* msDec, u64* algoBytes) {
  if (lastB == 0 || iters <= 0) throw Error(AMG_ERR_RANGE, "amg_bench_decode: no batch has been applied yet");
  const size_t B = lastB;
  hashTmp.ensure(ctx, B * 32 + 64);
  RawRows raw = rawRows();
  *algoBytes = (u64)lastBytes + 48ull * lastM + 8ull * lastP + 96ull * B;
#ifndef AMG_EMU
  cudaEvent_t e[4]; for (auto& x : e) cudaEventCreate(&x);
  const u8* batchArena = arena.p;
  cudaEventRecord(e[0], ctx.stream);
  for (int i = 0; i < iters; i++) sha_range(ctx, ShaTilesArgs{batchArena, chOff.p, chLen.p, hashTmp.p, errWord.p, nullptr, 0u, (u32)B}, false);
  cudaEventRecord(e[1], ctx.stream);
  for (int i = 0; i < iters; i++) runDecodeTiles(batchArena, B, lastBytes, deflList.p, lastDeflCount, lastDeflStart);
  cudaEventRecord(e[2], ctx.stream);
  for (int i = 0; i < iters; i++) {
    if (lastNumLarge > 0) foreach(ctx, (size_t)NCOLS * lastNumLarge, DecodeColumnKernel{batchArena, largeList.p, lastNumLarge, hot.p, nOps.p, nPreds.p, rawBase.p, rawPredBase.p, applied.p, raw, errWord.p, nullptr});
  }
  cudaEventRecord(e[3], ctx.stream);
  CUDA_CHECK(cudaEventSynchronize(e[3]));
  float a, b, c; cudaEventElapsedTime(&a, e[0], e[1]); cudaEventElapsedTime(&b, e[1], e[2]); cudaEventElapsedTime(&c, e[2], e[3]);
  *msSha = a / iters; *msParse = b / iters; *msDec = lastNumLarge > 0 ? c / iters : 0.f;
  for (auto& x : e) cudaEventDestroy(x);
#else
  *msSha = *msParse = *msDec = 0;
#endif
}

// Change history of a loaded document (reference new.js:1887-1912 computeHashGraph -> columnar.js:876-981): rebuilds every
// loaded change - ops from the document rows and the deletions implied by their succ lists, preds, actor tables, canonical
// column bytes - and its hash. Kernels: history.cuh. Nothing persistent is touched until the heads check has passed.
inline void Engine::computeHashGraph() {
  if (haveHashGraph) return;
  if (!unknownCols.empty()) throw Error(AMG_ERR_UNSUPPORTED, "amgpu: the change history of a loaded document that holds columns with unknown ids cannot be reconstructed");
  const size_t L = numLoaded, N = numRows, S = numSucc, A = actorIds.size();
  if (L == 0) { haveHashGraph = true; return; }
  if (L >= (1u << 29) || N + S >= (1u << 30)) throw Error(AMG_ERR_UNSUPPORTED, "amgpu: document too large for history reconstruction");
  DocRows d = doc.view();
  HostClock hclk; const bool htrace = getenv("AMG_PAR_DOC_TRACE") != nullptr;
  auto hmark = [&](const char* what) { if (htrace) { sync(ctx); fprintf(stderr, "amgpu history: %-26s %9.2f ms\n", what, hclk.ms()); } };
  dev_memset(ctx, errWord.p, 0, 16); errSnapLaunches = ~0ull;
  // ---- 1. change metadata columns (the same decoders as save() after load())
  auto loadedCol = [&](u32 id) -> const HostChange& { static const u32 IDS[9] = {0x01, 0x03, 0x13, 0x23, 0x35, 0x40, 0x43, 0x56, 0x57}; for (int k = 0; k < 9; k++) if (IDS[k] == id) return loadedCols[k]; return loadedCols[0]; };
  DBuf<long long> cActor, cSeq, cMaxOp, cTime, cDepsNum, cExtra, depIdxV, scratchV; DBuf<u32> msgOff, msgLen, extraOff, extraLen, tmpOff, tmpLen, depsNum32, depBase, depIdx;
  for (DBuf<long long>* b : {&cActor, &cSeq, &cMaxOp, &cTime, &cDepsNum, &cExtra, &scratchV}) b->ensure(ctx, L + 1);
  for (DBuf<u32>* b : {&msgOff, &msgLen, &extraOff, &extraLen, &tmpOff, &tmpLen, &depsNum32}) b->ensure(ctx, L + 2);
  depBase.ensure(ctx, L + 2);
  auto decodeCol = [&](int kind, u32 id, size_t count, long long* out, u32* so, u32* sl) {
    const HostChange& c = loadedCol(id);
    if (count >= parDocMinRows && c.len > 0) {
      const u8* bytes = arena.p + c.off; bool ok = false;
      if (kind == LC_UINT) ok = parCols.toI64(bytes, c.len, false, count, out);
      else if (kind == LC_DELTA) ok = parCols.deltaToI64(bytes, c.len, count, out);
      else if (kind == LC_EXTRA_LEN) ok = parCols.extraLenColumn(bytes, c.len, count, out, so, sl, loadedCol(0x57).off);
      if (ok) return;
    }
    if (count) foreach_warp(ctx, 1, LoadedColKernel{kind, arena.p, c.off, c.len, loadedCol(0x57).off, (u32)count, out, so, sl, nullptr});
  };
  decodeCol(LC_UINT, 0x01, L, cActor.p, tmpOff.p, tmpLen.p);
  decodeCol(LC_DELTA, 0x03, L, cSeq.p, tmpOff.p, tmpLen.p);
  decodeCol(LC_DELTA, 0x13, L, cMaxOp.p, tmpOff.p, tmpLen.p);
  decodeCol(LC_DELTA, 0x23, L, cTime.p, tmpOff.p, tmpLen.p);
  decodeCol(LC_STRING, 0x35, L, scratchV.p, msgOff.p, msgLen.p);
  decodeCol(LC_UINT, 0x40, L, cDepsNum.p, tmpOff.p, tmpLen.p);
  decodeCol(LC_EXTRA_LEN, 0x56, L, cExtra.p, extraOff.p, extraLen.p);
  foreach(ctx, L, HistI64ToU32Kernel{cDepsNum.p, depsNum32.p});
  scan_exclusive(ctx, scanTmp, depsNum32.p, depBase.p, L);
  const size_t D = readU32(depBase.p + L);
  depIdxV.ensure(ctx, D + 1); depIdx.ensure(ctx, D + 2);
  decodeCol(LC_DELTA, 0x43, D, depIdxV.p, tmpOff.p, tmpLen.p);
  if (D) foreach(ctx, D, HistI64ToU32Kernel{depIdxV.p, depIdx.p});
  // ---- 2. actor order (hex string order = byte order), representatives
  std::vector<u32> order(A), rankH(A), repOffH(A), repLenH(A);
  for (size_t a = 0; a < A; a++) { order[a] = (u32)a; repOffH[a] = actorRep[a].first; repLenH[a] = actorRep[a].second; }
  std::sort(order.begin(), order.end(), [&](u32 x, u32 y) { return actorIds[x] < actorIds[y]; });
  for (size_t i = 0; i < A; i++) rankH[order[i]] = (u32)i;
  DBuf<u32> rankD, actorOfRank, repOff, repLen;
  for (DBuf<u32>* b : {&rankD, &actorOfRank, &repOff, &repLen}) b->ensure(ctx, A + 1);
  h2d(ctx, rankD.p, rankH.data(), A * 4); h2d(ctx, actorOfRank.p, order.data(), A * 4); h2d(ctx, repOff.p, repOffH.data(), A * 4); h2d(ctx, repLen.p, repLenH.data(), A * 4);
  const int ctrBits = bits_for(maxOp + 1), idBits = std::min(64, ctrBits + 16);
  hmark("change columns decoded");
  // ---- 3. (successor, predecessor) pairs -> pred lists and deletions
  DBuf<u64> predKey, succKey, keyA, keyB, groupId, opId; DBuf<u32> pairRow, valA, pairRowSorted, head, groupIdx, groupStart, groupRow, isDel, delSlot, idRows;
  DBuf<u64> idSorted; idSorted.ensure(ctx, N + 1); idRows.ensure(ctx, N + 1);
  if (N) { foreach(ctx, N, HistIdKeyKernel{d, idSorted.p, idRows.p}); radix_sort_pairs(ctx, sortTmp, idSorted, idRows, N, 0, idBits); }
  size_t G = 0, numDel = 0;
  for (DBuf<u64>* b : {&predKey, &succKey, &keyA, &keyB}) b->ensure(ctx, S + 1);
  for (DBuf<u32>* b : {&pairRow, &valA, &pairRowSorted, &head, &groupIdx}) b->ensure(ctx, S + 2);
  if (S) {
    foreach(ctx, N, HistPairKernel{d, succOff.p, succ.p, rankD.p, predKey.p, succKey.p, pairRow.p});
    foreach(ctx, S, HistIotaKernel{valA.p});
    d2d(ctx, keyA.p, predKey.p, S * 8);
    radix_sort_pairs(ctx, sortTmp, keyA, valA, S, 0, idBits);                 // by predecessor (counter, actor order) ...
    foreach(ctx, S, HistGatherKeyKernel{succKey.p, valA.p, keyB.p});
    radix_sort_pairs(ctx, sortTmp, keyB, valA, S, 0, idBits);                 // ... then, stably, by successor id
    foreach(ctx, S, HistGatherU32Kernel{pairRow.p, valA.p, pairRowSorted.p});
    foreach(ctx, S, HistGroupHeadKernel{keyB.p, head.p});
    scan_exclusive(ctx, scanTmp, head.p, groupIdx.p, S);
    G = readU32(groupIdx.p + S);
  }
  for (DBuf<u32>* b : {&groupStart, &groupRow, &isDel, &delSlot}) b->ensure(ctx, G + 2);
  groupId.ensure(ctx, G + 1);
  if (G) {
    foreach(ctx, S, HistGroupKernel{head.p, groupIdx.p, keyB.p, (u32)S, idSorted.p, idRows.p, (u32)N, groupStart.p, groupId.p, groupRow.p, isDel.p});
    scan_exclusive(ctx, scanTmp, isDel.p, delSlot.p, G);
    numDel = readU32(delSlot.p + G);
  }
  const size_t M = N + numDel;
  DBuf<u32> opSrc, opPredStart, opPredNum, opOrder, opChange, predNumSorted, opPredBase;
  opId.ensure(ctx, M + 1); for (DBuf<u32>* b : {&opSrc, &opPredStart, &opPredNum, &opOrder, &opChange, &predNumSorted}) b->ensure(ctx, M + 2);
  opPredBase.ensure(ctx, M + 3);
  if (N) foreach(ctx, N, HistRowOpKernel{d, opId.p, opSrc.p, opPredStart.p, opPredNum.p});
  if (G) foreach(ctx, G, HistGroupOpKernel{groupStart.p, groupId.p, groupRow.p, isDel.p, delSlot.p, pairRowSorted.p, (u32)G, (u32)S, (u32)N, opId.p, opSrc.p, opPredStart.p, opPredNum.p});
  hmark("preds and deletions");
  // ---- 4. ops by (actor, counter); changes by (actor, seq); every op finds its change
  DBuf<u64> opKey, chKey; opKey.ensure(ctx, M + 1); chKey.ensure(ctx, L + 1);
  DBuf<u32> changeOrder, actorStart, chOpStart, chNOps; changeOrder.ensure(ctx, L + 1); actorStart.ensure(ctx, A + 2); chOpStart.ensure(ctx, L + 2); chNOps.ensure(ctx, L + 2);
  if (M) { foreach(ctx, M, HistOpKeyKernel{opId.p, opKey.p, opOrder.p}); radix_sort_pairs(ctx, sortTmp, opKey, opOrder, M, 0, ctrBits); radix_sort_pairs(ctx, sortTmp, opKey, opOrder, M, 48, 64); }
  foreach(ctx, L, HistChangeKeyKernel{cActor.p, cSeq.p, chKey.p, changeOrder.p});
  radix_sort_pairs(ctx, sortTmp, chKey, changeOrder, L, 0, 40); radix_sort_pairs(ctx, sortTmp, chKey, changeOrder, L, 40, 57);
  foreach(ctx, A + 1, HistLowerBoundKernel{chKey.p, (u32)L, 40, actorStart.p});
  dev_memset(ctx, chOpStart.p, 0, (L + 1) * 4); dev_memset(ctx, chNOps.p, 0, (L + 1) * 4);
  if (M) {
    foreach(ctx, M, HistAssignKernel{opKey.p, actorStart.p, changeOrder.p, cMaxOp.p, (u32)A, numApplied == L ? 1 : 0, opChange.p, errWord.p});
    foreach(ctx, M, HistChangeStartKernel{opChange.p, chOpStart.p});
    foreach(ctx, M, HistChangeCountKernel{opChange.p, chNOps.p});
    foreach(ctx, M, HistCheckIdsKernel{opKey.p, opChange.p, chOpStart.p, chNOps.p, cMaxOp.p, errWord.p});
    foreach(ctx, M, HistPredNumSortedKernel{opPredNum.p, opOrder.p, predNumSorted.p});
    scan_exclusive(ctx, scanTmp, predNumSorted.p, opPredBase.p, M);
  } else dev_memset(ctx, opPredBase.p, 0, 8);
  const size_t P = M ? readU32(opPredBase.p + M) : 0;
  checkErr(actorIds);
  hmark("ops assigned to changes");
  // ---- 5. the other actors of every change
  HistOpView view{d, opId.p, opSrc.p, opPredStart.p, opPredNum.p, opOrder.p, pairRowSorted.p, (u32)N};
  DBuf<u32> slotCnt, slotBase, uniq, uniqSlot, otherStart; DBuf<u64> slotKey, other; DBuf<u32> slotVal;
  slotCnt.ensure(ctx, M + 2); slotBase.ensure(ctx, M + 3); otherStart.ensure(ctx, L + 3);
  size_t Q = 0, U = 0;
  if (M) { foreach(ctx, M, HistActorSlotCountKernel{view, slotCnt.p}); scan_exclusive(ctx, scanTmp, slotCnt.p, slotBase.p, M); Q = readU32(slotBase.p + M); }
  slotKey.ensure(ctx, Q + 1); slotVal.ensure(ctx, Q + 1); uniq.ensure(ctx, Q + 2); uniqSlot.ensure(ctx, Q + 3);
  if (Q) {
    foreach(ctx, M, HistActorPairKernel{view, slotBase.p, opChange.p, cActor.p, rankD.p, slotKey.p});
    foreach(ctx, Q, HistIotaKernel{slotVal.p});
    radix_sort_pairs(ctx, sortTmp, slotKey, slotVal, Q, 0, 64);
    foreach(ctx, Q, HistUniqueKernel{slotKey.p, uniq.p});
    scan_exclusive(ctx, scanTmp, uniq.p, uniqSlot.p, Q);
    U = readU32(uniqSlot.p + Q);
  }
  other.ensure(ctx, U + 1);
  if (U) foreach(ctx, Q, HistOtherFillKernel{slotKey.p, uniq.p, uniqSlot.p, other.p});
  foreach(ctx, L + 1, HistLowerBoundKernel{other.p, (u32)U, 16, otherStart.p});
  hmark("actor tables");
  // ---- 6. local actor indexes and delta values, then the bytes (two passes)
  DBuf<u32> objA, keyAi, predA, outLen, outOff, depsAt, bodyAt, chOffD; DBuf<long long> keyDelta, predDelta;
  objA.ensure(ctx, M + 1); keyAi.ensure(ctx, M + 1); keyDelta.ensure(ctx, M + 1); predA.ensure(ctx, P + 1); predDelta.ensure(ctx, P + 1);
  for (DBuf<u32>* b : {&outLen, &depsAt, &bodyAt, &chOffD}) b->ensure(ctx, L + 2);
  outOff.ensure(ctx, L + 3);
  HistChanges hc{cActor.p, cSeq.p, cMaxOp.p, cTime.p, msgOff.p, msgLen.p, cDepsNum.p, extraOff.p, extraLen.p};
  foreach(ctx, L, HistPrepKernel{view, hc, chOpStart.p, chNOps.p, opPredBase.p, other.p, otherStart.p, rankD.p, objA.p, keyAi.p, keyDelta.p, predA.p, predDelta.p});
  HistEncodeKernel enc{0, view, hc, arena.p, chOpStart.p, chNOps.p, opPredBase.p, (u32)M, (u32)P, other.p, otherStart.p, repOff.p, repLen.p, actorOfRank.p,
                       objA.p, keyAi.p, keyDelta.p, predA.p, predDelta.p, outLen.p, outOff.p, nullptr, 0, depsAt.p, bodyAt.p};
  foreach(ctx, L, enc);
  scan_exclusive(ctx, scanTmp, outLen.p, outOff.p, L);
  excl64.ensure(ctx, L + 2); scan_exclusive64(ctx, scanTmp, ParColumnDecoder::PcDeltaInputU32{outLen.p}, excl64.p, L);
  u64 tot64 = 0; u32 lastLen = 0; d2h(ctx, &tot64, excl64.p + L - 1, 8); d2h(ctx, &lastLen, outLen.p + L - 1, 4); sync(ctx);
  const u64 T = tot64 + lastLen;
  if ((u64)arenaLen + T + 64 >= 0xfff00000ULL) throw Error(AMG_ERR_UNSUPPORTED, "amgpu: change arena limited to 4 GiB per document");
  arena.ensure(ctx, arenaLen + T + 64, arenaLen);
  enc.pass = 1; enc.arena = arena.p; enc.outArena = arena.p; enc.outBase = (u32)arenaLen;
  foreach(ctx, L, enc);
  foreach(ctx, L, HistChOffKernel{outOff.p, (u32)arenaLen, chOffD.p});
  checkErr(actorIds);
  hmark("changes encoded");
  // ---- 7. dependency levels (host: one pass over the dependency indexes), hashes level by level
  std::vector<u32> depsNumH(L), depBaseH(L + 1), depIdxH(D), level(L), list(L);
  d2h(ctx, depsNumH.data(), depsNum32.p, L * 4); d2h(ctx, depBaseH.data(), depBase.p, (L + 1) * 4); if (D) d2h(ctx, depIdxH.data(), depIdx.p, D * 4); sync(ctx);
  u32 maxLevel = 0; std::vector<u8> isDep(L, 0);
  for (size_t k = 0; k < L; k++) {
    u32 lv = 0;
    for (u32 i = 0; i < depsNumH[k]; i++) { const u32 di = depIdxH[depBaseH[k] + i]; if (di >= k) throw Error(AMG_ERR_RANGE, "No hash for index " + std::to_string(di) + " while processing index " + std::to_string(k)); lv = std::max(lv, level[di] + 1); isDep[di] = 1; }
    level[k] = lv; maxLevel = std::max(maxLevel, lv);
  }
  std::vector<u32> levelStart(maxLevel + 2, 0);
  for (size_t k = 0; k < L; k++) levelStart[level[k] + 1]++;
  for (u32 l = 0; l <= maxLevel; l++) levelStart[l + 1] += levelStart[l];
  { std::vector<u32> at(levelStart.begin(), levelStart.end() - 1); for (size_t k = 0; k < L; k++) list[at[level[k]]++] = (u32)k; }
  DBuf<u32> listD; listD.ensure(ctx, L + 1); h2d(ctx, listD.p, list.data(), L * 4);
  DBuf<u8> newHashes; newHashes.ensure(ctx, L * 32 + 64); d2d(ctx, newHashes.p, hashes.p, L * 32);   // scratch copy: committed only after the heads check
  {
    const HistHashKernel hk{listD.p, arena.p, chOffD.p, outLen.p, depsAt.p, bodyAt.p, cDepsNum.p, depBase.p, depIdx.p, (u32)L, newHashes.p, errWord.p};
    auto wide = [&](u32 l) { HistHashKernel k = hk; k.list = listD.p + levelStart[l]; const size_t cnt = levelStart[l + 1] - levelStart[l]; if (cnt) foreach(ctx, cnt, k); };
#ifdef AMG_EMU
    for (u32 l = 0; l <= maxLevel; l++) wide(l);
#else
    DBuf<u32> levelStartD; levelStartD.ensure(ctx, levelStart.size() + 1); h2d(ctx, levelStartD.p, levelStart.data(), levelStart.size() * 4);
    const u32 kNarrow = 1024;   // levels up to this many changes are walked by one CTA (k_hist_hash_chain); wider ones get their own launch
    for (u32 l = 0; l <= maxLevel;) {
      if (levelStart[l + 1] - levelStart[l] > kNarrow) { wide(l); l++; continue; }
      u32 r = l; while (r <= maxLevel && levelStart[r + 1] - levelStart[r] <= kNarrow) r++;
      k_hist_hash_chain<<<1, 256, 0, ctx.stream>>>(hk, levelStartD.p, l, r - l);
      CUDA_CHECK(cudaGetLastError()); ctx.launches++;
      l = r;
    }
    sync(ctx);   // levelStartD is a local
#endif
  }
  checkErr(actorIds);
  hmark("hashes (all levels)");
  // ---- 8. heads: the changes nobody depends on must be exactly the document's heads (columnar.js:968-980)
  {
    size_t nHeads = 0; for (size_t k = 0; k < L; k++) if (!isDep[k]) nHeads++;
    bool ok = numApplied != L || nHeads == heads.size();   // (changes applied after the load have moved the heads)
    std::vector<std::array<u8, 32>> got(heads.size());
    if (headIndexesUnknown) {   // loaded without head indexes: the heads are the changes nobody depends on, matched by hash
      if (numApplied != L) throw Error(AMG_ERR_INTERNAL, "amgpu: head indexes must be resolved right after the load");
      std::vector<u32> cand; for (size_t k = 0; k < L; k++) if (!isDep[k]) cand.push_back((u32)k);
      ok = cand.size() == heads.size();
      std::vector<std::array<u8, 32>> ch(cand.size());
      if (ok) { for (size_t i = 0; i < cand.size(); i++) d2h(ctx, ch[i].data(), newHashes.p + (size_t)cand[i] * 32, 32); sync(ctx); }
      for (size_t i = 0; i < heads.size() && ok; i++) {
        size_t j = 0; while (j < cand.size() && ch[j] != heads[i]) j++;
        if (j == cand.size()) ok = false; else headIdx[i] = cand[j];
      }
      if (ok) headIndexesUnknown = false;
    } else if (numApplied == L) {
      for (size_t i = 0; i < heads.size(); i++) d2h(ctx, got[i].data(), newHashes.p + (size_t)headIdx[i] * 32, 32);
      sync(ctx);
      for (size_t i = 0; i < heads.size() && ok; i++) if (isDep[headIdx[i]] || got[i] != heads[i]) ok = false;
    }
    if (!ok) throw Error(AMG_ERR_RANGE, "Mismatched heads hashes: the document's heads are not the hashes of its reconstructed changes");
  }
  hmark("heads checked");
  // ---- 9. commit: bytes into the arena and its host mirror, hashes, change table
  if (hostArena.size() == arenaLen) {   // the mirror is complete: keep it complete (otherwise it is fetched when asked for)
    hostArena.resize(arenaLen + T);
    if (T) d2h(ctx, hostArena.data() + arenaLen, arena.p + arenaLen, T);
  }
  d2d(ctx, hashes.p, newHashes.p, L * 32);
  std::vector<u32> offH(L), lenH(L); d2h(ctx, offH.data(), chOffD.p, L * 4); d2h(ctx, lenH.data(), outLen.p, L * 4); sync(ctx);
  for (size_t k = 0; k < L; k++) changes[k] = HostChange{offH[k], lenH[k]};
  arenaLen += T; haveHashGraph = true; historyRebuilt = L;
  hmark("committed");
}

// Parity hook: one document column through the parallel or the serial decoder (include/amgpu.h)
inline int Engine::debugDecodeColumn(const u8* bytes, size_t len, int kind, size_t n, bool parallel, long long* out) {
  if (kind < 0 || kind > 3 || len >= 0x7fffffffULL || n >= 0x7fffffffULL) throw Error(AMG_ERR_RANGE, "amg_debug_decode_column: bad arguments");
  DBuf<u8> colBytes; DBuf<long long> outD; DBuf<u32> tmp;
  colBytes.ensure(ctx, len + 64); dev_memset(ctx, colBytes.p, 0, len + 64); if (len) h2d(ctx, colBytes.p, bytes, len);
  outD.ensure(ctx, n + 1); tmp.ensure(ctx, n + 1);
  if (parallel) {
    bool ok = false;
    if (kind == 0) ok = parCols.toI64(colBytes.p, len, false, n, outD.p);
    else if (kind == 1) ok = parCols.toI64(colBytes.p, len, true, n, outD.p);
    else if (kind == 2) ok = parCols.deltaToI64(colBytes.p, len, n, outD.p);
    else { ok = parCols.boolean(colBytes.p, len, n, tmp.p); if (ok && n) foreach(ctx, n, U32ToI64Kernel{tmp.p, outD.p}); }
    if (!ok) { sync(ctx); return 1; }
  } else {
    dev_memset(ctx, errWord.p, 0, 16); errSnapLaunches = ~0ull;
    foreach_warp(ctx, 1, DebugColumnKernel{kind, colBytes.p, (u32)len, (u32)n, outD.p, tmp.p, errWord.p});
    checkErr(actorIds);
  }
  if (n) d2h(ctx, out, outD.p, n * 8);
  sync(ctx);
  return 0;
}

// Parity hook: hashes, op counts and the raw decoded columns (change-local values) of a batch, without touching the document.
inline void Engine::decodeRaw(const u8* blob, const u64* offsets, size_t n, u8* hashesOut, u32* nOpsOut, u32** rowsOut, size_t* totalOps, size_t* totalPreds) {
  std::vector<u32> off(n), len(n); std::string staged;
  for (size_t i = 0; i < n; i++) {
    const u8* p = blob + offsets[i]; const size_t l = offsets[i + 1] - offsets[i];
    std::string inflated; if (l > 8 && p[8] == 2) inflated = inflateChange(p, l);
    off[i] = (u32)staged.size(); len[i] = (u32)(inflated.empty() ? l : inflated.size());
    if (inflated.empty()) staged.append((const char*)p, l); else staged += inflated;
  }
  DBuf<u8> ar; ar.ensure(ctx, staged.size() + 64); h2d(ctx, ar.p, staged.data(), staged.size()); dev_memset(ctx, ar.p + staged.size(), 0, 64);
  chOff.ensure(ctx, n); chLen.ensure(ctx, n); h2d(ctx, chOff.p, off.data(), n * 4); h2d(ctx, chLen.p, len.data(), n * 4);
  dev_memset(ctx, errWord.p, 0, 16); errSnapLaunches = ~0ull; hashTmp.ensure(ctx, n * 32 + 64);
  foreach(ctx, n, ShaKernel{ar.p, chOff.p, chLen.p, hashTmp.p, errWord.p, nullptr, nullptr});
  applied.ensure(ctx, n); dev_memset(ctx, applied.p, 1, n);
  u32 tot[4] = {0, 0, 0, 0}; void* dst[4] = {&tot[0], &tot[1], &tot[2], &tot[3]};
  runDecodeTiles(ar.p, n, staged.size());
  readWords({{decTotalsPtr(), 4}, {decTotalsPtr() + 1, 4}, {decTotalsPtr() + 2, 4}, {decTotalsPtr() + 3, 4}}, dst);
  if (decodeOverflowed(tot)) { runDecodeTiles(ar.p, n, staged.size()); readWords({{decTotalsPtr(), 4}, {decTotalsPtr() + 1, 4}, {decTotalsPtr() + 2, 4}, {decTotalsPtr() + 3, 4}}, dst); }
  checkErr(actorIds);
  const size_t M = tot[0];
  DBuf<u32>* cols[12] = {&r_objActor, &r_objCtr, &r_keyActor, &r_keyCtr, &r_keyStrOff, &r_keyStrLen, &r_insert, &r_action, &r_valLen, &r_valOff, &r_predNum, &r_predOff};
  RawRows raw = rawRows();
  if (tot[3] & 1u) {
    largeFlag.ensure(ctx, n + 1); largeSlot.ensure(ctx, n + 2); largeList.ensure(ctx, n + 1);
    foreach(ctx, n, LargeFlagKernel{nOps.p, applied.p, largeFlag.p});
    scan_exclusive(ctx, scanTmp, largeFlag.p, largeSlot.p, n);
    const size_t nl = readU32(largeSlot.p + n);
    if (nl > 0) {
      foreach(ctx, n, CompactKernel{largeFlag.p, largeSlot.p, largeList.p});
      foreach(ctx, (size_t)NCOLS * nl, DecodeColumnKernel{ar.p, largeList.p, nl, hot.p, nOps.p, nPreds.p, rawBase.p, rawPredBase.p, applied.p, raw, errWord.p, nullptr});
    }
  }
  foreach(ctx, n, RaiseDecErrKernel{decErr.p, errWord.p});
  checkErr(actorIds);
  d2h(ctx, hashesOut, hashTmp.p, n * 32); d2h(ctx, nOpsOut, nOps.p, n * 4);
  const size_t P = tot[1];
  opBase.ensure(ctx, n + 1); predBase.ensure(ctx, n + 1);
  scan_exclusive(ctx, scanTmp, nOps.p, opBase.p, n); scan_exclusive(ctx, scanTmp, nPreds.p, predBase.p, n);
  DBuf<u32> gathered; gathered.ensure(ctx, 12 * (M + 1) + 2 * (P + 1));
  if (M) foreach(ctx, M, GatherRawKernel{n, opBase.p, predBase.p, rawBase.p, rawPredBase.p, raw, gathered.p, M, gathered.p + 12 * M, P});
  u32* rows = (u32*)malloc(sizeof(u32) * (12 * (M + 1) + 2 * (P + 1)));
  d2h(ctx, rows, gathered.p, (12 * M + 2 * P) * 4);
  if (totalPreds) *totalPreds = P;
  sync(ctx); *rowsOut = rows; *totalOps = M; lastB = 0;
}

}  // namespace amg

namespace amg {

extern "C" void amg_host_sha256(const uint8_t* data, size_t len, uint8_t out[32]);   // hostsha.cc (x86 SHA extensions when present)
inline void host_sha256(const u8* data, size_t len, u8 out[32]) { amg_host_sha256(data, len, out); }
inline std::string inflateRawBytes(const u8* p, size_t n) {
  z_stream zs; memset(&zs, 0, sizeof(zs));
  if (inflateInit2(&zs, -15) != Z_OK) throw Error(AMG_ERR_INTERNAL, "inflateInit failed");
  std::string out; out.resize(std::max<size_t>(n * 6, 1024)); zs.next_in = (Bytef*)p; zs.avail_in = (uInt)n; size_t produced = 0;
  while (true) {
    zs.next_out = (Bytef*)out.data() + produced; zs.avail_out = (uInt)(out.size() - produced);
    int rc = inflate(&zs, Z_NO_FLUSH); produced = out.size() - zs.avail_out;
    if (rc == Z_STREAM_END) break;
    if (rc != Z_OK && rc != Z_BUF_ERROR) { inflateEnd(&zs); throw Error(AMG_ERR_RANGE, "invalid deflate data"); }
    if (zs.avail_out == 0) out.resize(out.size() * 2); else if (zs.avail_in == 0) { inflateEnd(&zs); throw Error(AMG_ERR_RANGE, "unexpected end of deflate data"); }
  }
  inflateEnd(&zs); out.resize(produced); return out;
}

// Backend.save() (reference new.js:2033-2055, columnar.js:983-1004): change metadata columns (re-derived from the change
// headers that live in the arena) and the 16 document op columns (from the document table and its succ lists), encoded
// on the device (encode.cuh); the container (column directory, DEFLATE of columns >= 256 bytes, checksum) is assembled
// on the host, as the reference does.
inline void Engine::saveDocument(std::string& result) {
  if (!loadedDoc.empty()) { result = loadedDoc; return; }   // unchanged since Backend.load (new.js:2034)
  if (!encoder) encoder.reset(new ColumnEncoder(ctx, scanTmp));
  ColumnEncoder& enc = *encoder; enc.outLen = 0;
  HostClock sclk; const bool strace = getenv("AMG_PAR_DOC_TRACE") != nullptr;
  auto smark = [&](const char* what) { if (strace) { sync(ctx); fprintf(stderr, "amgpu save: %-28s %8.2f ms\n", what, sclk.ms()); } };
  struct Col { u32 id; size_t off, len; };
  std::vector<Col> changeCols, opCols;
  auto add = [&](std::vector<Col>& cols, u32 id, size_t len) { cols.push_back({id, enc.outLen - len, len}); };
  const size_t C = numApplied, N = numRows, S = numSucc, L = numLoaded, K = C - L;   // K changes have their bytes in the arena
  dev_memset(ctx, errWord.p, 0, 16); errSnapLaunches = ~0ull;
  saveVals.ensure(ctx, std::max(std::max(C, N), S) + 2);
  // ---- change metadata (new.js:1680-1692 appendChange); the first L rows come from the loaded document's own columns
  if (C > 0) {
    u32 totalDeps = 0, loadedDeps = 0;
    if (K > 0) {
      chPairs.ensure(ctx, K); chOff.ensure(ctx, K); chLen.ensure(ctx, K);
      h2d(ctx, chPairs.p, changes.data() + L, K * sizeof(HostChange));
      foreach(ctx, K, SplitPairsKernel{chPairs.p, chOff.p, chLen.p});
      meta.ensure(ctx, K); colOff.ensure(ctx, (size_t)NCOLS * K); colLen.ensure(ctx, (size_t)NCOLS * K);
      nOps.ensure(ctx, K + 1); nPreds.ensure(ctx, K + 1); nDeps.ensure(ctx, K + 1); nActors.ensure(ctx, K + 1);
      foreach(ctx, K, ParseKernel{arena.p, chOff.p, chLen.p, K, meta.p, colOff.p, colLen.p, nOps.p, nPreds.p, nDeps.p, nActors.p, errWord.p, flagWord.p + 8});
      depBase.ensure(ctx, K + 1); scan_exclusive(ctx, scanTmp, nDeps.p, depBase.p, K);
      totalDeps = readU32(depBase.p + K);
      depIdx.ensure(ctx, totalDeps + 1); primary.ensure(ctx, K);
      const size_t tcap = pow2_at_least(2 * C + 2);
      hashTable.ensure(ctx, tcap); dev_memset(ctx, hashTable.p, 0xff, tcap * 4);
      foreach(ctx, C, HashInsertKernel{hashes.p, hashTable.p, (u64)tcap - 1});
      foreach(ctx, K, ResolveDepsKernelT<ChangeMeta>{arena.p, hashes.p, hashTable.p, (u64)tcap - 1, meta.p, nDeps.p, L, depBase.p, depIdx.p, primary.p});
    }
    auto loadedCol = [&](u32 id) -> const HostChange& { static const u32 IDS[9] = {0x01, 0x03, 0x13, 0x23, 0x35, 0x40, 0x43, 0x56, 0x57}; for (int k = 0; k < 9; k++) if (IDS[k] == id) return loadedCols[k]; return loadedCols[0]; };
    if (L > 0) {   // number of dependency indexes the loaded changes carry
      DBuf<u64>& sumD = pairSucc; sumD.ensure(ctx, 1);
      const HostChange& dn = loadedCol(0x40);
      u64 sum = 0;
      if (!(L >= parDocMinRows && dn.len > 0 && parCols.sumColumn(arena.p + dn.off, dn.len, L, &sum))) {
        foreach_warp(ctx, 1, LoadedColKernel{LC_SUM, arena.p, dn.off, dn.len, 0, 0, nullptr, nullptr, nullptr, sumD.p});
        d2h(ctx, &sum, sumD.p, 8); sync(ctx);
      }
      loadedDeps = (u32)sum;
    }
    saveVals.ensure(ctx, std::max<size_t>(std::max(std::max(C, N), S), (size_t)loadedDeps + totalDeps) + 2);
    saveStrOff.ensure(ctx, std::max(C, N) + 1); saveStrLen.ensure(ctx, std::max(C, N) + 1);
    auto loadedVal = [&](int kind, u32 id, u32 count) {
      if (L == 0) return; const HostChange& c = loadedCol(id);
      if (count >= parDocMinRows && c.len > 0) {   // long history: the parallel column decoders (doccols.cuh); they decline what is not canonical
        const u8* bytes = arena.p + c.off; bool ok = false;
        if (kind == LC_UINT) ok = parCols.toI64(bytes, c.len, false, count, saveVals.p);
        else if (kind == LC_DELTA) ok = parCols.deltaToI64(bytes, c.len, count, saveVals.p);
        else if (kind == LC_EXTRA_LEN) ok = parCols.extraLenColumn(bytes, c.len, count, saveVals.p, saveStrOff.p, saveStrLen.p, loadedCol(0x57).off);
        if (ok) return;
      }
      foreach_warp(ctx, 1, LoadedColKernel{kind, arena.p, c.off, c.len, loadedCol(0x57).off, count, saveVals.p, saveStrOff.p, saveStrLen.p, nullptr});
    };
    auto changeVal = [&](int which) { if (K == 0) return; foreach(ctx, K, SaveChangeValKernel{which, arena.p, meta.p, actorSlots.p, (u64)actorCap - 1, saveVals.p + L, saveStrOff.p + L, saveStrLen.p + L, errWord.p}); };
    loadedVal(LC_UINT, 0x01, (u32)L);  changeVal(SM_ACTOR);     add(changeCols, 0x01, enc.rleNum(saveVals.p, C, false));
    loadedVal(LC_DELTA, 0x03, (u32)L); changeVal(SM_SEQ);       add(changeCols, 0x03, enc.deltaNum(saveVals.p, C));
    loadedVal(LC_DELTA, 0x13, (u32)L); changeVal(SM_MAX_OP);    add(changeCols, 0x13, enc.deltaNum(saveVals.p, C));
    loadedVal(LC_DELTA, 0x23, (u32)L); changeVal(SM_TIME);      add(changeCols, 0x23, enc.deltaNum(saveVals.p, C));
    loadedVal(LC_STRING, 0x35, (u32)L); if (K > 0) foreach(ctx, K, SaveMessageKernel{meta.p, saveStrOff.p + L, saveStrLen.p + L});
                                       add(changeCols, 0x35, enc.rle(StrCol{arena.p, saveStrOff.p, saveStrLen.p}, C));
    loadedVal(LC_UINT, 0x40, (u32)L);  changeVal(SM_DEPS_NUM);  add(changeCols, 0x40, enc.rleNum(saveVals.p, C, false));
    loadedVal(LC_DELTA, 0x43, loadedDeps); if (totalDeps > 0) foreach(ctx, totalDeps, SaveDepIndexKernel{depIdx.p, saveVals.p + loadedDeps});
                                       add(changeCols, 0x43, enc.deltaNum(saveVals.p, (size_t)loadedDeps + totalDeps));
    loadedVal(LC_EXTRA_LEN, 0x56, (u32)L); changeVal(SM_EXTRA_LEN); add(changeCols, 0x56, enc.rleNum(saveVals.p, C, false));
                                       add(changeCols, 0x57, enc.raw(arena.p, saveStrOff.p, saveStrLen.p, C));
    checkErr(actorIds);
  }
  // ---- document ops (columnar.js:60-82)
  if (N > 0) {
    DocRows d = doc.view();
    saveStrOff.ensure(ctx, N + 1); saveStrLen.ensure(ctx, N + 1);
    auto opVal = [&](int which) { foreach(ctx, N, SaveOpValKernel{which, d, succOff.p, saveVals.p}); };
    opVal(SC_OBJ_ACTOR); add(opCols, 0x01, enc.rleNum(saveVals.p, N, false));
    opVal(SC_OBJ_CTR);   add(opCols, 0x02, enc.rleNum(saveVals.p, N, false));
    opVal(SC_KEY_ACTOR); add(opCols, 0x11, enc.rleNum(saveVals.p, N, false));
    opVal(SC_KEY_CTR);   add(opCols, 0x13, enc.deltaNum(saveVals.p, N));
                         add(opCols, 0x15, enc.rle(StrCol{arena.p, d.keyStrOff, d.keyStrLen}, N));
    opVal(SC_ID_ACTOR);  add(opCols, 0x21, enc.rleNum(saveVals.p, N, false));
    opVal(SC_ID_CTR);    add(opCols, 0x23, enc.deltaNum(saveVals.p, N));
    foreach(ctx, N, SaveInsertKernel{d, saveStrLen.p});
                         add(opCols, 0x34, enc.boolean(saveStrLen.p, N));
    opVal(SC_ACTION);    add(opCols, 0x42, enc.rleNum(saveVals.p, N, false));
    opVal(SC_VAL_LEN);   add(opCols, 0x56, enc.rleNum(saveVals.p, N, false));
    foreach(ctx, N, SaveValBytesKernel{d, saveStrLen.p});
                         add(opCols, 0x57, enc.raw(arena.p, d.valOff, saveStrLen.p, N));
    // chldActor 0x61 / chldCtr 0x63: always null in this format version -> empty
    opVal(SC_SUCC_NUM);  add(opCols, 0x80, enc.rleNum(saveVals.p, N, false));
    if (S > 0) {
      foreach(ctx, S, SaveSuccValKernel{0, succ.p, saveVals.p}); add(opCols, 0x81, enc.rleNum(saveVals.p, S, false));
      foreach(ctx, S, SaveSuccValKernel{1, succ.p, saveVals.p}); add(opCols, 0x83, enc.deltaNum(saveVals.p, S));
    }
  }
  smark("columns encoded (device)");
  std::vector<u8> raw(enc.outLen);
  if (enc.outLen) { d2h(ctx, raw.data(), enc.out.p, enc.outLen); sync(ctx); }
  // ---- host: DEFLATE of large columns (columnar.js:1052-1057), directory, container (columnar.js:659-686)
  struct Packed { u32 id; std::string data; };
  std::vector<Packed> packed; std::vector<size_t> firstOp;
  auto pack = [&](const std::vector<Col>& cols) { for (auto& c : cols) if (c.len > 0) packed.push_back({c.id, std::string((const char*)raw.data() + c.off, c.len)}); };
  pack(changeCols); const size_t numChangeCols = packed.size(); pack(opCols);
  if (!unknownCols.empty() && N > 0) {   // columns with ids this version does not know: host-encoded from the values kept per op (unknowncols.hpp)
    std::vector<std::pair<u32, std::string>> extra; appendUnknownDocColumns(extra);
    for (auto& e : extra) if (!e.second.empty()) packed.push_back({e.first, e.second});
    std::stable_sort(packed.begin() + numChangeCols, packed.end(), [](const Packed& a, const Packed& b) { return (a.id & ~8u) < (b.id & ~8u); });
  }
  {
    std::vector<std::thread> ts; std::vector<std::string> errs(packed.size());
    for (size_t k = 0; k < packed.size(); k++) if (packed[k].data.size() >= 256) ts.emplace_back([&, k] {
      z_stream zs; memset(&zs, 0, sizeof(zs));
      if (deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { errs[k] = "deflateInit failed"; return; }
      std::string comp; comp.resize(deflateBound(&zs, (uLong)packed[k].data.size()));
      zs.next_in = (Bytef*)packed[k].data.data(); zs.avail_in = (uInt)packed[k].data.size(); zs.next_out = (Bytef*)comp.data(); zs.avail_out = (uInt)comp.size();
      const int rc = ::deflate(&zs, Z_FINISH); comp.resize(zs.total_out); deflateEnd(&zs);
      if (rc != Z_STREAM_END) { errs[k] = "deflate failed"; return; }
      packed[k].data.swap(comp); packed[k].id |= 8;
    });
    for (auto& t : ts) t.join();
    for (auto& e : errs) if (!e.empty()) throw Error(AMG_ERR_INTERNAL, e);
  }
  smark("columns deflated (host)");
  std::string body;
  auto uleb = [&](u64 v) { do { u8 b = v & 0x7f; v >>= 7; if (v) b |= 0x80; body.push_back((char)b); } while (v); };
  uleb(actorIds.size()); for (auto& a : actorIds) { uleb(a.size()); body += a; }
  uleb(heads.size()); for (auto& h : heads) body.append((const char*)h.data(), 32);
  uleb(numChangeCols); for (size_t k = 0; k < numChangeCols; k++) { uleb(packed[k].id); uleb(packed[k].data.size()); }
  uleb(packed.size() - numChangeCols); for (size_t k = numChangeCols; k < packed.size(); k++) { uleb(packed[k].id); uleb(packed[k].data.size()); }
  for (auto& p : packed) body += p.data;
  for (u32 i : headIdx) uleb(i);
  std::string head; head.push_back(0); { u64 v = body.size(); do { u8 b = v & 0x7f; v >>= 7; if (v) b |= 0x80; head.push_back((char)b); } while (v); }
  std::string hashed = head + body; u8 digest[32]; host_sha256((const u8*)hashed.data(), hashed.size(), digest);
  smark("container assembled + hashed");
  static const u8 magic[4] = {0x85, 0x6f, 0x4a, 0x83};
  result.assign((const char*)magic, 4); result.append((const char*)digest, 4); result += hashed;
}

// Backend.load(data) = new BackendDoc(buffer) (reference new.js:1709-1750): the document chunk's op columns are already in
// document order with succ lists, so loading = container check + column decode + one finalize pass. The container
// checksum (one SHA-256 over the whole chunk: inherently serial) and the DEFLATE of large columns are host pre-passes,
// as in SURVEY.md §2 row 12; column expansion and everything downstream run on the device.
inline void Engine::loadDocument(const u8* buf, size_t len) {
  if (numApplied != 0 || numRows != 0) throw Error(AMG_ERR_INTERNAL, "load needs a fresh backend");
  HostClock lclk; const bool ltrace = getenv("AMG_PAR_DOC_TRACE") != nullptr;
  auto lmark = [&](const char* what) { if (ltrace) { sync(ctx); fprintf(stderr, "amgpu load: %-28s %8.2f ms\n", what, lclk.ms()); } };
  // columnar.js:688-708 decodeContainerHeader
  if (len < 10 || buf[0] != 0x85 || buf[1] != 0x6f || buf[2] != 0x4a || buf[3] != 0x83) throw Error(AMG_ERR_RANGE, "Data does not begin with magic bytes 85 6f 4a 83");
  ByteReader r(buf, 8, (u32)len); const u32 chunkType = buf[8]; r.pos = 9; const u64 chunkLen = r.uleb();
  if (r.err || (u64)r.pos + chunkLen > len) throw Error(AMG_ERR_RANGE, "buffer ended with incomplete number");
  u8 digest[32]; host_sha256(buf + 8, r.pos + (size_t)chunkLen - 8, digest);
  if (memcmp(digest, buf + 4, 4) != 0) throw Error(AMG_ERR_RANGE, "checksum does not match data");
  if ((u64)r.pos + chunkLen != len) throw Error(AMG_ERR_RANGE, "Encoded document has trailing data");
  if (chunkType != 0) throw Error(AMG_ERR_RANGE, "Unexpected chunk type: " + std::to_string(chunkType));
  lmark("container checksum");
  // columnar.js:1006-1038 decodeDocumentHeader
  std::vector<std::string> actors; const u64 numActors = r.uleb();
  for (u64 i = 0; i < numActors && !r.err; i++) { const u64 l = r.uleb(); if ((u64)r.pos + l > len) { r.err = KE_SUBARRAY; break; } actors.emplace_back((const char*)buf + r.pos, l); r.skip(l); }
  std::vector<std::array<u8, 32>> hs; const u64 numHeads = r.uleb();
  for (u64 i = 0; i < numHeads && !r.err; i++) { if ((u64)r.pos + 32 > len) { r.err = KE_SUBARRAY; break; } std::array<u8, 32> h; memcpy(h.data(), buf + r.pos, 32); hs.push_back(h); r.skip(32); }
  struct ColInfo { u32 id; u64 len; std::string data; };
  auto readInfo = [&](std::vector<ColInfo>& cols) {
    const u64 n = r.uleb(); long long last = -1;
    for (u64 i = 0; i < n && !r.err; i++) { const u64 id = r.uleb(), l = r.uleb(); if (last >= 0 && ((u32)id & ~8u) <= ((u32)last & ~8u)) throw Error(AMG_ERR_RANGE, "Columns must be in ascending order"); last = (long long)id; cols.push_back({(u32)id, l, std::string()}); }
  };
  std::vector<ColInfo> changeCols, opCols; readInfo(changeCols); readInfo(opCols);
  std::vector<std::pair<ColInfo*, const u8*>> deflated;
  auto readData = [&](std::vector<ColInfo>& cols) {
    for (auto& c : cols) {
      if (r.err || (u64)r.pos + c.len > len) throw Error(AMG_ERR_RANGE, "subarray exceeds buffer size");
      if (c.id & 8) deflated.emplace_back(&c, buf + r.pos); else c.data.assign((const char*)buf + r.pos, (size_t)c.len);
      r.skip(c.len);
    }
  };
  readData(changeCols); readData(opCols);
  {   // DEFLATEd columns (columnar.js:1022-1027): independent streams, one host thread each when there are several large ones
    std::vector<std::string> errs(deflated.size()); std::vector<int> codes(deflated.size(), 0); std::vector<std::thread> ts;
    auto one = [&](size_t k) { try { ColInfo& c = *deflated[k].first; c.data = inflateRawBytes(deflated[k].second, (size_t)c.len); c.id ^= 8; } catch (Error& e) { errs[k] = e.what(); codes[k] = e.code; } catch (std::exception& e) { errs[k] = e.what(); codes[k] = AMG_ERR_INTERNAL; } };
    for (size_t k = 0; k < deflated.size(); k++) { if (deflated.size() > 1 && deflated[k].first->len >= (64u << 10)) ts.emplace_back(one, k); else one(k); }
    for (auto& t : ts) t.join();
    for (size_t k = 0; k < errs.size(); k++) if (codes[k]) throw Error(codes[k], errs[k]);   // the first in column order, as a sequential reader would meet it
  }
  lmark("columns inflated");
  if (r.err) throw Error(AMG_ERR_RANGE, "buffer ended with incomplete number");
  std::vector<u32> headsIndexes; if (!r.done()) for (u64 i = 0; i < numHeads; i++) headsIndexes.push_back((u32)r.uleb());
  if (actors.size() > 65535) throw Error(AMG_ERR_UNSUPPORTED, "amgpu: more than 65535 actors in one document");
  // ---- stage actor ids and op columns in the arena
  static const u32 DOC_IDS[16] = {0x01, 0x02, 0x11, 0x13, 0x15, 0x21, 0x23, 0x34, 0x42, 0x56, 0x57, 0x61, 0x63, 0x80, 0x81, 0x83};
  DocCols dc; memset(&dc, 0, sizeof(dc));
  hostArena.resize(0); std::vector<std::pair<u32, u32>> reps;
  { size_t total = 0; for (auto& a : actors) total += a.size(); for (auto& c : opCols) total += c.data.size(); for (auto& c : changeCols) total += c.data.size(); hostArena.reserve(total + 64); }   // one pinned allocation, not one per append
  for (auto& a : actors) { reps.emplace_back((u32)hostArena.size(), (u32)a.size()); hostArena.append(a.data(), a.size()); }
  for (auto& c : opCols) for (int k = 0; k < 16; k++) if (c.id == DOC_IDS[k]) { dc.off[k] = (u32)hostArena.size(); dc.len[k] = (u32)c.data.size(); hostArena.append(c.data.data(), c.data.size()); }
  {   // the change metadata columns stay available for a later save() (new.js:1717 keeps them as encoders)
    static const u32 CHANGE_IDS[9] = {0x01, 0x03, 0x13, 0x23, 0x35, 0x40, 0x43, 0x56, 0x57};
    for (int k = 0; k < 9; k++) loadedCols[k] = HostChange{(u32)hostArena.size(), 0};
    for (auto& c : changeCols) for (int k = 0; k < 9; k++) if (c.id == CHANGE_IDS[k]) { loadedCols[k] = HostChange{(u32)hostArena.size(), (u32)c.data.size()}; hostArena.append(c.data.data(), c.data.size()); }
  }
  const size_t cur = hostArena.size();
  if (cur + 64 >= 0xfff00000ULL) throw Error(AMG_ERR_UNSUPPORTED, "amgpu: change arena limited to 4 GiB per document");
  arena.ensure(ctx, cur + 64); h2d(ctx, arena.p, hostArena.data(), cur); dev_memset(ctx, arena.p + cur, 0, 64);
  lmark("staged + uploaded");
  // ---- change metadata: clock (new.js:1645-1675 readDocumentChanges). A long history is decoded and checked on the device
  // (doccols.cuh + one stable sort by actor); a short one, or one the device path declines (malformed columns, a sequence
  // error to report), by the same readers on the host, which produce the reference's error messages.
  std::vector<u64> clk(actors.size(), 0); size_t numChanges = 0;
  auto hostClock = [&]()
  {
    const std::string* actorCol = nullptr; const std::string* seqCol = nullptr;
    for (auto& c : changeCols) { if (c.id == 0x01) actorCol = &c.data; if (c.id == 0x03) seqCol = &c.data; }
    static const std::string empty;
    const std::string& ac = actorCol ? *actorCol : empty; const std::string& sc = seqCol ? *seqCol : empty;
    RleReader ar((const u8*)ac.data(), 0, (u32)ac.size(), 0), sr((const u8*)sc.data(), 0, (u32)sc.size(), 1); long long seqAcc = 0;
    while (!ar.done()) {
      long long a = 0, d = 0; u32 o, l; const bool an = ar.next(a, o, l), sn = sr.next(d, o, l);
      if (ar.r.err || sr.r.err) throw Error(AMG_ERR_RANGE, "malformed change metadata columns");
      if (!an || (u64)a >= actors.size()) throw Error(AMG_ERR_RANGE, "actor index out of range");
      if (sn) seqAcc += d;
      const u64 seq = sn ? (u64)seqAcc : 0;
      if (seq != 1 && seq != clk[a] + 1) throw Error(AMG_ERR_RANGE, "Expected seq " + std::to_string(clk[a] + 1) + ", got " + std::to_string(seq) + " for actor " + hex_of((const u8*)actors[a].data(), actors[a].size()));
      clk[a] = seq; numChanges++;
    }
  };
  {
    bool onDevice = false; const HostChange ac = loadedCols[0], sc = loadedCols[1]; u32 total = 0;
    if (ac.len >= parDocMinRows / 8 + 16 && sc.len > 0 && parCols.rleRecords(arena.p + ac.off, ac.len, &total) && total >= parDocMinRows) {
      const size_t n = total; DBuf<long long> aV, sV; aV.ensure(ctx, n + 1); sV.ensure(ctx, n + 1);
      if (parCols.toI64(arena.p + ac.off, ac.len, false, n, aV.p) && parCols.deltaToI64(arena.p + sc.off, sc.len, n, sV.p)) {
        sortKeys.ensure(ctx, n + 1); sortVals.ensure(ctx, n + 1); DBuf<u64> clkD; clkD.ensure(ctx, actors.size() + 1); dev_memset(ctx, clkD.p, 0, (actors.size() + 1) * 8);
        dev_memset(ctx, flagWord.p, 0, 16);
        foreach(ctx, n, ClockKeyKernel{aV.p, (u32)actors.size(), sortKeys.p, sortVals.p, flagWord.p});
        sortPairs(sortKeys, sortVals, n, bits_for(actors.size() > 1 ? actors.size() - 1 : 1));
        foreach(ctx, n, ClockCheckKernel{sortKeys.p, sortVals.p, sV.p, (u32)n, clkD.p, flagWord.p});
        u32 bad = 0; d2h(ctx, &bad, flagWord.p, 4); if (!actors.empty()) d2h(ctx, clk.data(), clkD.p, actors.size() * 8); sync(ctx);
        if (!bad) { onDevice = true; numChanges = n; } else std::fill(clk.begin(), clk.end(), 0);
      }
    }
    if (!onDevice) hostClock();
  }
  lmark("clock");
  if (!headsIndexes.empty() && headsIndexes.size() != hs.size()) headsIndexes.clear();
  // several heads without indexes (new.js:1734-1737: the hashes are known, their change indexes are not): the indexes are
  // found by reconstructing the change history right after the load (computeHashGraph below)
  bool headIdxUnknown = false;
  if (headsIndexes.empty()) { if (hs.size() == 1) headsIndexes.push_back((u32)(numChanges ? numChanges - 1 : 0)); else if (!hs.empty()) { headIdxUnknown = true; headsIndexes.assign(hs.size(), 0xffffffffu); } }
  dev_memset(ctx, errWord.p, 0, 16); errSnapLaunches = ~0ull; dev_memset(ctx, flagWord.p, 0, 16);
  // Number of rows = values of the action column, number of succ entries = sum of succNum. Long columns take the parallel
  // decoder (doccols.cuh); short, malformed or non-canonical ones the serial walkers, which also report the errors.
  { const char* e = getenv("AMG_PAR_DOC_MIN"); if (e) parDocMinRows = (size_t)strtoull(e, nullptr, 10); }
  size_t N = 0, S = 0; u32 serialMask = 0xffffu; bool counted = false;
  auto colBytes = [&](int k) { return arena.p + dc.off[k]; };
  auto ensureRows = [&]() {
    for (DBuf<u32>* b : {&r_objActor, &r_objCtr, &r_keyActor, &r_keyCtr, &r_keyStrOff, &r_keyStrLen, &r_insert, &r_action, &r_valLen, &r_valOff, &r_predNum, &r_predOff, &o_change, &o_time}) b->ensure(ctx, N + 1);
  };
  if (dc.len[8] >= parDocMinRows / 8 + 16) {   // (a long column can still be a handful of records; then the serial count is instant anyway)
    u32 total = 0;
    if (parCols.rleRecords(colBytes(8), dc.len[8], &total) && total >= parDocMinRows && total < (1u << 29)) {
      N = total; ensureRows();
      u64 sum = 0;
      if (dc.len[13] == 0) { counted = true; S = 0; }
      else if (parCols.countColumn(colBytes(13), dc.len[13], N, r_predNum.p, r_predOff.p, &sum)) { counted = true; S = (size_t)sum; serialMask &= ~(1u << 13); }
    }
  }
  if (!counted) {   // short action column: rows by the serial record walk, the succ total still in parallel if there are many rows
    serialMask = 0xffffu;
    foreach(ctx, 1, DocCountRowsKernel{arena.p, dc, flagWord.p, errWord.p});
    u32 n32 = 0; d2h(ctx, &n32, flagWord.p, 4); sync(ctx); checkErr(actors);
    if (n32 >= parDocMinRows && n32 < (1u << 29)) {
      N = n32; ensureRows(); u64 sum = 0;
      if (dc.len[13] == 0) { counted = true; S = 0; }
      else if (parCols.countColumn(colBytes(13), dc.len[13], N, r_predNum.p, r_predOff.p, &sum)) { counted = true; S = (size_t)sum; serialMask &= ~(1u << 13); }
    }
  }
  if (!counted) {
    serialMask = 0xffffu;
    foreach(ctx, 1, DocCountKernel{arena.p, dc, flagWord.p, errWord.p});
    u32 cnt[2]; d2h(ctx, cnt, flagWord.p, 8); sync(ctx); checkErr(actors);
    N = cnt[0]; S = cnt[1];
  }
  lmark("rows counted");
  if (N >= (1u << 29)) throw Error(AMG_ERR_UNSUPPORTED, "amgpu: more than 2^29 document rows");
  ensureRows();
  r_predActor.ensure(ctx, S + 1); r_predCtr.ensure(ctx, S + 1);
  RawRows raw{r_objActor.p, r_objCtr.p, r_keyActor.p, r_keyCtr.p, r_keyStrOff.p, r_keyStrLen.p, r_insert.p, r_action.p, r_valLen.p, r_valOff.p, r_predNum.p, r_predOff.p, r_predActor.p, r_predCtr.p};
  if (N >= parDocMinRows) {
    struct Plan { int k, col; u32* out; };   // document column -> row field, as DocColumnKernel maps them
    const Plan plan[] = {{0, CX_OBJ_ACTOR, raw.objActor}, {1, CX_OBJ_CTR, raw.objCtr}, {2, CX_KEY_ACTOR, raw.keyActor}, {3, CX_KEY_CTR, raw.keyCtr}, {4, CX_KEY_STR, nullptr},
                         {5, CX_OBJ_ACTOR, o_change.p}, {6, CX_KEY_CTR, o_time.p}, {7, CX_INSERT, raw.insert}, {8, CX_ACTION, raw.action}, {9, CX_VAL_LEN, raw.valLen},
                         {13, CX_PRED_NUM, raw.predNum}, {14, CX_PRED_ACTOR, raw.predActor}, {15, CX_PRED_CTR, raw.predCtr}};
    for (const Plan& pl : plan) {
      if (!((serialMask >> pl.k) & 1u)) continue;
      const size_t cnt = (pl.k == 14 || pl.k == 15) ? S : N; const u32 len = dc.len[pl.k]; const u8* bytes = colBytes(pl.k); bool done = false;
      if (len == 0) {
        RawRows rr = raw; if (pl.k == 5) rr.objActor = o_change.p; if (pl.k == 6) rr.keyCtr = o_time.p;
        if (cnt > 0) foreach(ctx, cnt, DocAbsentKernel{pl.col, rr});
        done = true;
      } else if (cnt > 0) switch (pl.col) {
        case CX_OBJ_ACTOR: case CX_OBJ_CTR: case CX_KEY_ACTOR: case CX_ACTION: case CX_PRED_ACTOR: done = parCols.toU32(bytes, len, cnt, pl.out); break;
        case CX_KEY_CTR: case CX_PRED_CTR: done = parCols.deltaToU32(bytes, len, cnt, pl.out); break;
        case CX_INSERT: done = parCols.boolean(bytes, len, cnt, pl.out); break;
        case CX_VAL_LEN: { u64 sum = 0; done = parCols.lenColumn(bytes, len, cnt, raw.valLen, raw.valOff, dc.off[10], &sum) && sum <= dc.len[10]; } break;
        case CX_PRED_NUM: { u64 sum = 0; done = parCols.countColumn(bytes, len, cnt, raw.predNum, raw.predOff, &sum) && sum == S; } break;
        default: break;   // utf8 keys: below
      }
      if (done) serialMask &= ~(1u << pl.k);
    }
    if ((serialMask >> 4) & 1u) {   // keyStr: serial over records, parallel over rows
      DBuf<u32>& recStart = parCols.recOff; DBuf<u32>& recStrOff = parCols.recTok; DBuf<u32>& recStrLen = parCols.recN;
      recStart.ensure(ctx, N + 3); recStrOff.ensure(ctx, N + 3); recStrLen.ensure(ctx, N + 3);
      foreach_warp(ctx, 1, DocKeyStrRecordsKernel{arena.p, dc.off[4], dc.len[4], (u32)N, recStart.p, recStrOff.p, recStrLen.p, flagWord.p, errWord.p});
      u32 rc[2]; d2h(ctx, rc, flagWord.p, 8); sync(ctx); checkErr(actors);
      foreach(ctx, N, DocKeyStrExpandKernel{recStart.p, recStrOff.p, recStrLen.p, rc[0], raw.keyStrOff, raw.keyStrLen});
      serialMask &= ~(1u << 4);
    }
  }
  if (getenv("AMG_PAR_DOC_TRACE")) fprintf(stderr, "amgpu load: %zu rows, %zu succ entries, columns left to the serial decoder: mask %04x (counted in parallel: %d)\n", N, S, serialMask & 0xe3ffu, counted ? 1 : 0);
  if (serialMask & 0xe3ffu) foreach_warp(ctx, 16, DocColumnKernel{arena.p, dc, (u32)N, (u32)S, raw, o_change.p, o_time.p, errWord.p, serialMask});
  lmark("columns decoded");
  doc.ensure(ctx, N + 1); succOff.ensure(ctx, N + 2); succ.ensure(ctx, S + 1);
  DBuf<u64>& maxOpD = pairSucc; maxOpD.ensure(ctx, 1); dev_memset(ctx, maxOpD.p, 0, 8);
  foreach(ctx, N, DocFinalizeKernel{raw, o_change.p, o_time.p, (u32)actors.size(), doc.view(), succOff.p, succ.p, maxOpD.p, errWord.p});
  { const u32 s32 = (u32)S; h2d(ctx, succOff.p + N, &s32, 4); }
  u64 mx = 0; d2h(ctx, &mx, maxOpD.p, 8); sync(ctx); checkErr(actors);
  lmark("rows finalized");
  {   // op columns with unknown ids: their values are kept per op (host; unknowncols.hpp) so that save() writes them again
    bool anyUnknown = false; for (auto& c : opCols) if (!is_known_doc_column(c.id)) anyUnknown = true;
    unknownCols.clear();
    if (anyUnknown && N > 0) {
      std::string all; std::vector<std::array<u32, 3>> cols;
      for (auto& c : opCols) { cols.push_back({c.id, (u32)all.size(), (u32)c.data.size()}); all += c.data; }
      std::vector<u64> ids(N); d2h(ctx, ids.data(), doc.id.p, N * 8); sync(ctx);
      const u32 e = read_unknown_columns((const u8*)all.data(), cols, N, is_known_doc_column, [&](size_t i, UnknownRow& row) {
        if (row.empty()) return;
        for (auto& kv : row) unknownCols.colIds.insert(kv.first);
        unknownCols.byOp[ids[i]] = row;
      });
      if (e == KE_UNSUPPORTED_OP) throw Error(AMG_ERR_RANGE, "unexpected VALUE_RAW column");
      if (e) throwKernelError((u64)e, actors);
    }
  }
  // ---- change history placeholders: only the head hashes are known (new.js:1727-1739)
  hashes.ensure(ctx, numChanges * 32 + 64); dev_memset(ctx, hashes.p, 0, numChanges * 32 + 64);
  if (!headIdxUnknown) for (size_t i = 0; i < hs.size(); i++) { if (headsIndexes[i] >= numChanges) throw Error(AMG_ERR_RANGE, "head index out of range"); h2d(ctx, hashes.p + (size_t)headsIndexes[i] * 32, hs[i].data(), 32); }
  sync(ctx);
  numRows = N; numSucc = S; numApplied = numChanges; arenaLen = cur; maxOp = mx; actorIds = actors; actorRep = reps; clock = clk;
  heads = hs; headIdx = headsIndexes;
  { std::vector<size_t> o(heads.size()); for (size_t i = 0; i < o.size(); i++) o[i] = i; std::sort(o.begin(), o.end(), [&](size_t a, size_t b) { return hs[a] < hs[b]; });
    for (size_t i = 0; i < o.size(); i++) { heads[i] = hs[o[i]]; headIdx[i] = headsIndexes[o[i]]; } }
  changes.assign(numChanges, HostChange{0, 0}); haveHashGraph = false; loadedDoc.assign((const char*)buf, len); numLoaded = numChanges; headIndexesUnknown = headIdxUnknown;
  lmark("host state");
  while (actorCap < 2 * (actorIds.size() + 16)) actorCap *= 2;
  actorSlots.ensure(ctx, actorCap); rebuildActorTable();
  if (headIndexesUnknown) computeHashGraph();   // finds the heads' change indexes (and leaves the history rebuilt)
}

}  // namespace amg
