// amgpu — extern "C" surface (include/amgpu.h) over amg::Engine, plus the host-side hash graph
// (reference backend/new.js:1921-2028: getChanges / getChangesAdded / getChangeByHash / getMissingDeps).
#include <unordered_map>
#include "../../include/amgpu.h"
#include "engine_impl.cuh"

using namespace amg;

struct amg_patch { const u8* p; size_t len; };   // view into the engine's pinned patch buffer
struct amg_buffers { std::vector<std::string> items; };

namespace {
typedef std::array<u8, 32> Hash;
struct HashHasher { size_t operator()(const Hash& h) const { size_t v; memcpy(&v, h.data(), sizeof(v)); return v; } };

// Host hash graph, filled lazily from the applied changes' headers (the reference defers it too: new.js:1887-1912)
struct HostGraph {
  size_t known = 0;
  std::vector<Hash> hash; std::vector<std::string> actor; std::vector<u64> seq; std::vector<std::vector<Hash>> deps;
  std::unordered_map<Hash, u32, HashHasher> indexByHash; std::unordered_map<Hash, std::vector<Hash>, HashHasher> dependents;
  std::map<std::string, std::vector<Hash>> hashesByActor;
};
}  // namespace

struct amg_backend {
  Engine eng; HostGraph g;
  explicit amg_backend(int dev) : eng(dev) {}

  void ensureGraph() {
    Engine& e = eng;
    if (!e.haveHashGraph) e.computeHashGraph();   // new.js:1922, 1980, 2000, 2015
    if (g.known == e.numApplied) return;
    const size_t from = g.known, to = e.numApplied;
    e.ensureHostMirror();   // the change headers are read from the host copy of the arena (fetched now if the batch came from pinned / device memory)
    std::vector<u8> hs((to - from) * 32); d2h(e.ctx, hs.data(), e.hashes.p + from * 32, hs.size()); sync(e.ctx);
    for (size_t i = from; i < to; i++) {
      Hash h; memcpy(h.data(), hs.data() + (i - from) * 32, 32);
      const HostChange& c = e.changes[i];
      ByteReader r(e.hostArena.data(), c.off + 8, c.off + c.len); r.pos++; r.uleb();
      const u64 nd = r.uleb(); std::vector<Hash> deps(nd);
      for (u64 k = 0; k < nd; k++) { memcpy(deps[k].data(), e.hostArena.data() + r.pos, 32); r.skip(32); }
      const u64 al = r.uleb(); std::string actor((const char*)e.hostArena.data() + r.pos, al); r.skip(al);
      const u64 seq = r.uleb();
      g.hash.push_back(h); g.actor.push_back(actor); g.seq.push_back(seq); g.deps.push_back(deps);
      g.indexByHash[h] = (u32)i; g.dependents[h];
      for (auto& d : deps) g.dependents[d].push_back(h);
      auto& v = g.hashesByActor[actor]; if (v.size() < seq) v.resize(seq); v[seq - 1] = h;
    }
    g.known = to;
  }
  // columnar.js:798-811 deflateChange: magic + checksum of the plain form, chunk type 2, raw DEFLATE of the body
  static std::string deflateChange(const std::string& plain) {
    ByteReader r((const u8*)plain.data(), 9, (u32)plain.size()); const u64 bodyLen = r.uleb();
    if (r.err || r.pos + bodyLen != plain.size()) throw amg::Error(AMG_ERR_INTERNAL, "deflateChange: malformed change");
    z_stream zs; memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw amg::Error(AMG_ERR_INTERNAL, "deflateInit failed");
    std::string comp; comp.resize(deflateBound(&zs, (uLong)bodyLen));
    zs.next_in = (Bytef*)plain.data() + r.pos; zs.avail_in = (uInt)bodyLen; zs.next_out = (Bytef*)comp.data(); zs.avail_out = (uInt)comp.size();
    const int rc = ::deflate(&zs, Z_FINISH); comp.resize(zs.total_out); deflateEnd(&zs);
    if (rc != Z_STREAM_END) throw amg::Error(AMG_ERR_INTERNAL, "deflate failed");
    std::string out(plain, 0, 8); out.push_back(2);
    { u64 v = comp.size(); do { u8 b = v & 0x7f; v >>= 7; if (v) b |= 0x80; out.push_back((char)b); } while (v); }
    return out + comp;
  }
  std::string changeBytes(u32 idx) {
    eng.ensureHostMirror();
    if (const HostChange* o = eng.originalOf(idx)) return std::string((const char*)eng.hostArena.data() + o->off, o->len);
    const HostChange& c = eng.changes[idx];
    std::string plain((const char*)eng.hostArena.data() + c.off, c.len);
    if (idx < eng.historyRebuilt && plain.size() >= 256) return deflateChange(plain);   // what encodeChange returns for a rebuilt change (columnar.js:738)
    return plain;
  }
};

namespace {
void setErr(amg_error* err, int code, const std::string& msg) {
  if (!err) return;
  err->code = code; snprintf(err->msg, sizeof(err->msg), "%s", msg.c_str());
}
#define AMG_GUARD(...) \
  try { __VA_ARGS__ } catch (amg::Error& e) { amg::drop_pending_peeks(); setErr(err, e.code, e.what()); return e.code; } \
  catch (std::exception& e) { amg::drop_pending_peeks(); setErr(err, AMG_INTERNAL_ERROR, e.what()); return AMG_INTERNAL_ERROR; }

amg_patch* serialize(const PatchOut& p) { return new amg_patch{p.bytes, p.bytesLen}; }
Hash toHash(const u8* p) { Hash h; memcpy(h.data(), p, 32); return h; }
std::string hashHex(const Hash& h) { return hex_of(h.data(), 32); }
}  // namespace

extern "C" {

amg_backend* amg_init(int cuda_device, amg_error* err) {
  try { return new amg_backend(cuda_device); }
  catch (amg::Error& e) { setErr(err, e.code, e.what()); return nullptr; }
  catch (std::exception& e) { setErr(err, AMG_INTERNAL_ERROR, e.what()); return nullptr; }
}
void amg_free(amg_backend* b) { delete b; }
amg_backend* amg_load(int cuda_device, const uint8_t* data, size_t len, amg_error* err) {
  amg_backend* b = nullptr;
  try { b = new amg_backend(cuda_device); b->eng.loadDocument(data, len); return b; }
  catch (amg::Error& e) { setErr(err, e.code, e.what()); delete b; return nullptr; }
  catch (std::exception& e) { setErr(err, AMG_INTERNAL_ERROR, e.what()); delete b; return nullptr; }
}
int amg_reset(amg_backend* b, amg_error* err) { AMG_GUARD(b->g = HostGraph(); b->eng.reset(); return 0;) }
int amg_reserve(amg_backend* b, size_t arena_bytes, amg_error* err) { AMG_GUARD(b->eng.hostArena.reserve(arena_bytes); b->eng.arena.ensure(b->eng.ctx, arena_bytes + 64, b->eng.arenaLen); return 0;) }

amg_backend* amg_clone(amg_backend* src, amg_error* err) {
  try {
    auto* b = new amg_backend(src->eng.ctx.device);
    Engine& d = b->eng; Engine& s = src->eng; Ctx& c = d.ctx;
    sync(s.ctx); s.ensureHostMirror();
    d.hostArena.assign(s.hostArena); d.arenaLen = s.arenaLen; d.arena.ensure(c, s.arenaLen + 64); d2d(c, d.arena.p, s.arena.p, s.arenaLen);
    d.numApplied = s.numApplied; d.hashes.ensure(c, s.numApplied * 32 + 64); d2d(c, d.hashes.p, s.hashes.p, s.numApplied * 32);
    d.numRows = s.numRows; d.doc.ensure(c, s.numRows + 1);
    d2d(c, d.doc.id.p, s.doc.id.p, s.numRows * 8); d2d(c, d.doc.obj.p, s.doc.obj.p, s.numRows * 8); d2d(c, d.doc.key.p, s.doc.key.p, s.numRows * 8);
    d2d(c, d.doc.keyStrOff.p, s.doc.keyStrOff.p, s.numRows * 4); d2d(c, d.doc.keyStrLen.p, s.doc.keyStrLen.p, s.numRows * 4); d2d(c, d.doc.flags.p, s.doc.flags.p, s.numRows * 4);
    d2d(c, d.doc.valLen.p, s.doc.valLen.p, s.numRows * 4); d2d(c, d.doc.valOff.p, s.doc.valOff.p, s.numRows * 4); d2d(c, d.doc.time.p, s.doc.time.p, s.numRows * 4);
    d.numSucc = s.numSucc; d.succOff.ensure(c, s.numRows + 2); d2d(c, d.succOff.p, s.succOff.p, (s.numRows + 1) * 4); d.succ.ensure(c, s.numSucc + 1); d2d(c, d.succ.p, s.succ.p, s.numSucc * 8);
    d.actorIds = s.actorIds; d.actorRep = s.actorRep; d.clock = s.clock; d.heads = s.heads; d.headIdx = s.headIdx; d.changes = s.changes; d.deflatedOriginal = s.deflatedOriginal; d.loadedDoc = s.loadedDoc; d.numLoaded = s.numLoaded; for (int k = 0; k < 9; k++) d.loadedCols[k] = s.loadedCols[k];
    d.unknownCols = s.unknownCols; d.queue = s.queue; d.queueOriginal = s.queueOriginal; d.maxOp = s.maxOp; d.haveHashGraph = s.haveHashGraph; d.historyRebuilt = s.historyRebuilt;
    while (d.actorCap < 2 * (d.actorIds.size() + 16)) d.actorCap *= 2;
    d.actorSlots.ensure(c, d.actorCap); d.rebuildActorTable();
    sync(c);
    return b;
  } catch (amg::Error& e) { amg::drop_pending_peeks(); setErr(err, e.code, e.what()); return nullptr; }
  catch (std::exception& e) { amg::drop_pending_peeks(); setErr(err, AMG_INTERNAL_ERROR, e.what()); return nullptr; }
}

int amg_apply_changes(amg_backend* b, const uint8_t* const* bufs, const size_t* lens, size_t n, int is_local, int want_patch, amg_patch** out, amg_error* err) {
  AMG_GUARD(PatchOut p; b->eng.applyChanges(bufs, lens, n, nullptr, nullptr, is_local != 0, want_patch != 0, p);
            if (out) *out = want_patch ? serialize(p) : nullptr; return 0;)
}
int amg_apply_changes_packed(amg_backend* b, const uint8_t* blob, const uint64_t* offsets, size_t n, int is_local, int want_patch, amg_patch** out, amg_error* err) {
  AMG_GUARD(amg::HostClock whole; { PatchOut p; b->eng.applyChanges(nullptr, nullptr, n, blob, (const u64*)offsets, is_local != 0, want_patch != 0, p);
            if (out) *out = want_patch ? serialize(p) : nullptr; } b->eng.lastPhaseMs[23] = whole.ms(); return 0;)   // [23]: the whole call as the ABI sees it
}
int amg_get_patch(amg_backend* b, amg_patch** out, amg_error* err) {
  AMG_GUARD(PatchOut p; b->eng.getPatch(p); *out = serialize(p); return 0;)
}
int amg_get_state(amg_backend* b, amg_patch** out, amg_error* err) {
  AMG_GUARD(PatchOut p; b->eng.fillPatchHeader(p); b->eng.finishPatch(p); *out = serialize(p); return 0;)
}
const uint8_t* amg_patch_bytes(const amg_patch* p, size_t* len) { *len = p->len; return p->p; }
void amg_patch_free(amg_patch* p) { delete p; }
const uint8_t* amg_arena(amg_backend* b, size_t* len) {
  try { b->eng.ensureHostMirror(); } catch (...) { amg::drop_pending_peeks(); *len = 0; return nullptr; }
  *len = b->eng.hostArena.size(); return b->eng.hostArena.data();
}

size_t amg_buffers_count(const amg_buffers* l) { return l->items.size(); }
const uint8_t* amg_buffers_get(const amg_buffers* l, size_t i, size_t* len) { *len = l->items[i].size(); return (const uint8_t*)l->items[i].data(); }
void amg_buffers_free(amg_buffers* l) { delete l; }
void amg_free_mem(void* p) { free(p); }

int amg_get_heads(amg_backend* b, amg_buffers** out, amg_error* err) {
  AMG_GUARD(auto* l = new amg_buffers(); for (auto& h : b->eng.heads) l->items.emplace_back((const char*)h.data(), 32); *out = l; return 0;)
}

// new.js:2033-2055
int amg_save(amg_backend* b, amg_buffers** out, amg_error* err) {
  AMG_GUARD(auto* l = new amg_buffers(); std::unique_ptr<amg_buffers> guard(l); l->items.emplace_back(); b->eng.saveDocument(l->items.back()); *out = guard.release(); return 0;)
}

// new.js:1921-1973
int amg_get_changes(amg_backend* b, const uint8_t* have_deps, size_t n, amg_buffers** out, amg_error* err) {
  AMG_GUARD(
    b->ensureGraph(); HostGraph& g = b->g; auto* l = new amg_buffers(); std::unique_ptr<amg_buffers> guard(l);
    if (n == 0) { for (size_t i = 0; i < b->eng.changes.size(); i++) l->items.push_back(b->changeBytes((u32)i)); *out = guard.release(); return 0; }
    std::vector<Hash> stack, toReturn; std::unordered_map<Hash, bool, HashHasher> seen;
    for (size_t i = 0; i < n; i++) {
      Hash h = toHash(have_deps + 32 * i); seen[h] = true;
      auto it = g.dependents.find(h); if (it == g.dependents.end() || !g.indexByHash.count(h)) throw amg::Error(AMG_RANGE_ERROR, "hash not found: " + hashHex(h));
      stack.insert(stack.end(), it->second.begin(), it->second.end());
    }
    // Reference quirk reproduced on purpose (new.js:1938-1955): the traversal stops at a change with an unseen dependency, but
    // the test below only looks at the stack and the heads - when that change was the last one on the stack and the heads
    // have all been seen, the fast path still answers, without the changes that are concurrent to `haveDeps`.
    while (!stack.empty()) {
      Hash h = stack.back(); stack.pop_back(); seen[h] = true; toReturn.push_back(h);
      bool all = true; for (auto& d : g.deps[g.indexByHash[h]]) if (!seen.count(d)) all = false;
      if (!all) break;
      auto& ds = g.dependents[h]; stack.insert(stack.end(), ds.begin(), ds.end());
    }
    bool headsSeen = true; for (auto& h : b->eng.heads) if (!seen.count(h)) headsSeen = false;
    if (stack.empty() && headsSeen) { for (auto& h : toReturn) l->items.push_back(b->changeBytes(g.indexByHash[h])); *out = guard.release(); return 0; }
    stack.clear(); for (size_t i = 0; i < n; i++) stack.push_back(toHash(have_deps + 32 * i)); seen.clear();
    while (!stack.empty()) {
      Hash h = stack.back(); stack.pop_back();
      if (!seen.count(h)) {
        auto it = g.indexByHash.find(h); if (it == g.indexByHash.end()) throw amg::Error(AMG_RANGE_ERROR, "hash not found: " + hashHex(h));
        auto& ds = g.deps[it->second]; stack.insert(stack.end(), ds.begin(), ds.end()); seen[h] = true;
      }
    }
    for (size_t i = 0; i < b->eng.changes.size(); i++) if (!seen.count(g.hash[i])) l->items.push_back(b->changeBytes((u32)i));
    *out = guard.release(); return 0;)
}
// new.js:1979-1997
int amg_get_changes_added(amg_backend* bn, amg_backend* bo, amg_buffers** out, amg_error* err) {
  AMG_GUARD(
    bn->ensureGraph(); bo->ensureGraph(); HostGraph& g = bn->g; auto* l = new amg_buffers();
    std::vector<Hash> stack = bn->eng.heads, toReturn; std::unordered_map<Hash, bool, HashHasher> seen;
    while (!stack.empty()) {
      Hash h = stack.back(); stack.pop_back();
      if (!seen.count(h) && !bo->g.indexByHash.count(h)) { seen[h] = true; toReturn.push_back(h); auto& ds = g.deps[g.indexByHash[h]]; stack.insert(stack.end(), ds.begin(), ds.end()); }
    }
    for (auto it = toReturn.rbegin(); it != toReturn.rend(); ++it) l->items.push_back(bn->changeBytes(g.indexByHash[*it]));
    *out = l; return 0;)
}
int amg_get_change_by_hash(amg_backend* b, const uint8_t hash[32], amg_buffers** out, amg_error* err) {
  AMG_GUARD(b->ensureGraph(); auto* l = new amg_buffers(); auto it = b->g.indexByHash.find(toHash(hash));
            if (it != b->g.indexByHash.end()) l->items.push_back(b->changeBytes(it->second)); *out = l; return 0;)
}
// new.js:2014-2028
int amg_get_missing_deps(amg_backend* b, const uint8_t* heads, size_t n, amg_buffers** out, amg_error* err) {
  AMG_GUARD(
    b->ensureGraph(); std::map<Hash, bool> allDeps, inQueue; Engine& e = b->eng;
    for (size_t i = 0; i < n; i++) allDeps[toHash(heads + 32 * i)] = true;
    e.ensureHostMirror();
    for (auto& q : e.queue) {
      Hash h;
      // hash of the queued change: SHA-256 over bytes [8..) — computed on the host for the (short) queue
      ByteReader r(e.hostArena.data(), q.off + 8, q.off + q.len); r.pos++; r.uleb();
      const u64 nd = r.uleb(); for (u64 k = 0; k < nd; k++) { allDeps[toHash(e.hostArena.data() + r.pos)] = true; r.skip(32); }
      u32 hh[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
      const u8* m = e.hostArena.data() + q.off + 8; const u32 mlen = q.len - 8; std::vector<u8> padded(m, m + mlen); padded.push_back(0x80);
      while (padded.size() % 64 != 56) padded.push_back(0); for (int i = 7; i >= 0; i--) padded.push_back((u8)(((u64)mlen * 8) >> (8 * i)));
      for (size_t o = 0; o < padded.size(); o += 64) { u32 w[16]; for (int i = 0; i < 16; i++) w[i] = (u32)padded[o + 4 * i] << 24 | (u32)padded[o + 4 * i + 1] << 16 | (u32)padded[o + 4 * i + 2] << 8 | padded[o + 4 * i + 3]; sha256_compress(hh, w, SHA_K); }
      for (int i = 0; i < 8; i++) { h[4 * i] = hh[i] >> 24; h[4 * i + 1] = hh[i] >> 16; h[4 * i + 2] = hh[i] >> 8; h[4 * i + 3] = hh[i]; }
      inQueue[h] = true;
    }
    auto* l = new amg_buffers();
    for (auto& kv : allDeps) if (!b->g.indexByHash.count(kv.first) && !inQueue.count(kv.first)) l->items.emplace_back((const char*)kv.first.data(), 32);
    *out = l; return 0;)
}
int amg_clock_of(amg_backend* b, const uint8_t* actor, size_t actor_len, uint64_t* seq_out, amg_error* err) {
  AMG_GUARD(std::string a((const char*)actor, actor_len); *seq_out = 0; Engine& e = b->eng;
            for (size_t i = 0; i < e.actorIds.size(); i++) if (e.actorIds[i] == a) *seq_out = e.clock[i]; return 0;)
}
int amg_hash_by_actor(amg_backend* b, const uint8_t* actor, size_t actor_len, uint64_t index, uint8_t hash_out[32], int* found, amg_error* err) {
  AMG_GUARD(b->ensureGraph(); *found = 0; auto it = b->g.hashesByActor.find(std::string((const char*)actor, actor_len));
            if (it != b->g.hashesByActor.end() && index < it->second.size()) { memcpy(hash_out, it->second[index].data(), 32); *found = 1; } return 0;)
}

int amg_last_timings(amg_backend* b, float* ms_out, int n) { for (int i = 0; i < n && i < 24; i++) ms_out[i] = b->eng.lastPhaseMs[i]; return 0; }
size_t amg_debug_marks(amg_backend* b, char* buf, size_t cap) {
  std::string s; for (auto& m : b->eng.dbgMarks) { char t[96]; snprintf(t, sizeof t, "%s=%.3f ", m.first, m.second); s += t; }
  if (cap) { snprintf(buf, cap, "%s", s.c_str()); } return s.size();
}
uint64_t amg_kernel_launches(amg_backend* b) { return b->eng.ctx.launches; }

int amg_debug_dump_ops(amg_backend* b, uint64_t** rows_out, size_t* n, uint64_t** succ_out, size_t* m, amg_error* err) {
  AMG_GUARD(
    Engine& e = b->eng; const size_t N = e.numRows, S = e.numSucc;
    std::vector<u64> id(N), obj(N), key(N), succ(S); std::vector<u32> flags(N), soff(N + 1), ksl(N);
    d2h(e.ctx, id.data(), e.doc.id.p, N * 8); d2h(e.ctx, obj.data(), e.doc.obj.p, N * 8); d2h(e.ctx, key.data(), e.doc.key.p, N * 8);
    d2h(e.ctx, flags.data(), e.doc.flags.p, N * 4); d2h(e.ctx, ksl.data(), e.doc.keyStrLen.p, N * 4); d2h(e.ctx, soff.data(), e.succOff.p, (N + 1) * 4); d2h(e.ctx, succ.data(), e.succ.p, S * 8); sync(e.ctx);
    u64* r = (u64*)malloc(sizeof(u64) * 8 * (N + 1)); u64* s = (u64*)malloc(sizeof(u64) * 2 * (S + 1));
    for (size_t i = 0; i < N; i++) {
      u64* o = r + 8 * i; const u64 none = ~0ULL;
      o[0] = obj[i] ? id_ctr(obj[i]) : none; o[1] = obj[i] ? id_actor(obj[i]) : none; o[2] = id_ctr(id[i]); o[3] = id_actor(id[i]);
      const bool list = ksl[i] == NULL32;
      o[4] = list ? id_ctr(key[i]) : none; o[5] = (list && key[i]) ? id_actor(key[i]) : none; o[6] = flags[i]; o[7] = soff[i + 1] - soff[i];
    }
    for (size_t i = 0; i < S; i++) { s[2 * i] = id_ctr(succ[i]); s[2 * i + 1] = id_actor(succ[i]); }
    *rows_out = (uint64_t*)r; *n = N; *succ_out = (uint64_t*)s; *m = S; return 0;)
}

int amg_debug_decode(amg_backend* b, const uint8_t* blob, const uint64_t* offsets, size_t n, uint8_t* hashes_out, uint32_t* n_ops_out, uint32_t** rows_out, size_t* total_ops, size_t* total_preds, amg_error* err) {
  AMG_GUARD(b->eng.decodeRaw(blob, (const u64*)offsets, n, hashes_out, n_ops_out, rows_out, total_ops, total_preds); return 0;)
}
int amg_debug_decode_column(amg_backend* b, const uint8_t* bytes, size_t len, int kind, size_t n, int parallel, int64_t* out, amg_error* err) {
  AMG_GUARD(return b->eng.debugDecodeColumn(bytes, len, kind, n, parallel != 0, (long long*)out);)
}
int amg_bench_decode(amg_backend* b, int iters, float* ms_sha, float* ms_parse, float* ms_decode, uint64_t* algo_bytes, amg_error* err) {
  AMG_GUARD(u64 bytes = 0; b->eng.benchDecode(iters, ms_sha, ms_parse, ms_decode, &bytes); *algo_bytes = bytes; return 0;)
}
}  // extern "C"
