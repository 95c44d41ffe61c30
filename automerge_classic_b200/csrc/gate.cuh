// amgpu — kernels #2: causal gate and actor interning.
//
// Replaces (reference paths relative to /root/reference):
//   backend/new.js:1550-1597  applyChanges(): duplicate-hash skip, dependency readiness, seq == clock+1,
//                             heads maintenance
//   backend/new.js:1822-1841  the retry loop over the queue of not-yet-ready changes
//   backend/new.js:1434-1451  getActorTable(): first-appearance actor numbering, change-local -> doc actor index
//
// Change hashes live in an open-addressing table keyed by the first 8 bytes of the SHA-256 digest
// (full 32-byte compare on hit). The reference applies, per pass over the queue, every change whose
// deps are all applied already (or earlier in the same pass); its application order is therefore
// (pass number, queue position). pass[c] = max over deps d of (pos[d] < pos[c] ? pass[d] : pass[d]+1)
// is computed as a monotone fixpoint (RelaxKernel, iterated until no change) — one round when the
// batch is already in causal order.
#pragma once
#include "decode.cuh"

namespace amg {

static const u32 EMPTY32 = 0xffffffffu;
static const u32 DEP_MISSING = 0xffffffffu;
static const u32 PASS_INF = 0x7fffffffu;

HD bool hash_eq(const u8* a, const u8* b) {
  const u64* x = reinterpret_cast<const u64*>(a); const u64* y = reinterpret_cast<const u64*>(b);
  return x[0] == y[0] && x[1] == y[1] && x[2] == y[2] && x[3] == y[3];
}
HD u64 hash_key(const u8* h) { return *reinterpret_cast<const u64*>(h); }
HD u64 load_u64_unaligned(const u8* p) { u64 v = 0; for (int i = 7; i >= 0; i--) v = (v << 8) | p[i]; return v; }

// inserts change g (global index) into the table; equal hashes keep the smallest index
struct HashInsertKernel {
  const u8* hashes /* [total][32] */; u32* table; u64 mask;
  HD void operator()(size_t g) const {
    const u8* h = hashes + g * 32; u64 slot = mix64(hash_key(h)) & mask;
    { const u64* x = reinterpret_cast<const u64*>(h); if ((x[0] | x[1] | x[2] | x[3]) == 0) return; }   // placeholder of a loaded change whose hash is not known
    while (true) {
      u32 cur = atomic_cas(&table[slot], EMPTY32, (u32)g);
      if (cur == EMPTY32) return;
      if (hash_eq(hashes + (size_t)cur * 32, h)) { atomic_min(&table[slot], (u32)g); return; }
      slot = (slot + 1) & mask;
    }
  }
};
HD u32 hash_lookup(const u8* hashes, const u32* table, u64 mask, const u8* h32 /* 8-byte aligned copy */) {
  u64 slot = mix64(hash_key(h32)) & mask;
  while (true) {
    const u32 cur = table[slot];
    if (cur == EMPTY32) return DEP_MISSING;
    if (hash_eq(hashes + (size_t)cur * 32, h32)) return cur;
    slot = (slot + 1) & mask;
  }
}

// per batch change: primary[b] = smallest global index with the same hash; resolves dependency hashes
template <class Meta> struct ResolveDepsKernelT {   // Meta: ChangeHot (apply) or ChangeMeta (save)
  const u8* arena; const u8* hashes; const u32* table; u64 mask; const Meta* meta; const u32* nDeps; size_t numApplied;
  const u32* depBase; u32* depIdx; u32* primary;
  HD void operator()(size_t b) const {
    const size_t g = numApplied + b;
    primary[b] = hash_lookup(hashes, table, mask, hashes + g * 32);
    const u32 n = nDeps[b], base = depBase[b];
    for (u32 j = 0; j < n; j++) {
      u64 tmp[4]; const u8* src = arena + meta[b].depsOff + 32 * j;
      for (int k = 0; k < 4; k++) tmp[k] = load_u64_unaligned(src + 8 * k);
      depIdx[base + j] = hash_lookup(hashes, table, mask, reinterpret_cast<const u8*>(tmp));
    }
  }
};

typedef ResolveDepsKernelT<ChangeHot> ResolveDepsKernel;

// one relaxation sweep; pass[b] starts at 1 for candidates, PASS_INF for changes that can never apply
struct RelaxKernel {
  const u32* depBase; const u32* depIdx; const u32* nDeps; const u32* primary; size_t numApplied; u32* pass; u32* changed; u32 maxPass;
  HD void operator()(size_t b) const {
    const size_t g = numApplied + b;
    if (primary[b] != (u32)g) return;   // duplicate of an applied change or of an earlier batch entry
    u32 p = 1; const u32 n = nDeps[b], base = depBase[b];
    for (u32 j = 0; j < n; j++) {
      const u32 d = depIdx[base + j];
      if (d == DEP_MISSING) { p = PASS_INF; break; }
      if (d < numApplied) continue;
      const u32 db = d - (u32)numApplied; const u32 pd = pass[db];
      if (pd >= PASS_INF) { p = PASS_INF; break; }
      const u32 need = db < b ? pd : pd + 1;
      if (need > p) p = need;
    }
    if (p > maxPass) p = PASS_INF;
    if (p != pass[b]) { pass[b] = p; *changed = 1; }
  }
};

// ---- the same change more than once among the entries that are not applied yet (a copy in the batch, another in the queue):
// the reference walks the queue in order pass after pass and applies whichever copy is ready first; the others are then
// duplicates (new.js:1556-1557). Every copy therefore takes part: best[group] = min over the copies of (pass, position),
// a dependency is satisfied by the best copy of its hash. Only run when such copies exist (flag from GateDupFlagKernel).
struct GateDupFlagKernel { const u32* primary; size_t numApplied; u32* flag; HD void operator()(size_t b) const { if (primary[b] >= numApplied && primary[b] != (u32)(numApplied + b)) *flag = 1; } };
struct GateBestKernel {
  const u32* primary; const u32* pass; size_t numApplied; u64* best;
  HD void operator()(size_t b) const { if (primary[b] < numApplied) return; atomic_min(&best[primary[b] - numApplied], ((u64)pass[b] << 32) | (u64)b); }
};
struct RelaxCopiesKernel {
  const u32* depBase; const u32* depIdx; const u32* nDeps; const u32* primary; size_t numApplied; const u64* best; u32* pass; u32* changed; u32 maxPass;
  HD void operator()(size_t b) const {
    if (primary[b] < numApplied) return;   // a copy of an applied change
    u32 p = 1; const u32 n = nDeps[b], base = depBase[b];
    for (u32 j = 0; j < n; j++) {
      const u32 d = depIdx[base + j];
      if (d == DEP_MISSING) { p = PASS_INF; break; }
      if (d < numApplied) continue;
      const u64 bd = best[d - (u32)numApplied]; const u32 pd = (u32)(bd >> 32), posd = (u32)bd;
      if (pd >= PASS_INF) { p = PASS_INF; break; }
      const u32 need = posd < b ? pd : pd + 1;
      if (need > p) p = need;
    }
    if (p > maxPass) p = PASS_INF;
    if (p != pass[b]) { pass[b] = p; *changed = 1; }
  }
};
// the best copy of every hash becomes its primary; dependencies point at it
struct GateWinnerKernel {
  const u64* best; size_t numApplied; u32* primary;
  HD void operator()(size_t b) const { if (primary[b] < numApplied) return; primary[b] = (u32)numApplied + (u32)best[primary[b] - numApplied]; }
};
struct GateDepWinnerKernel {
  const u64* best; size_t numApplied; u32* depIdx;
  HD void operator()(size_t j) const { const u32 d = depIdx[j]; if (d == DEP_MISSING || d < numApplied) return; depIdx[j] = (u32)numApplied + (u32)best[d - (u32)numApplied]; }
};

// ---------------------------------------------------------------- actor interning
// Byte-string table: slot = {hash64 of the bytes, min(firstSeen<<32 | change)}; identity is the 64-bit
// FNV-1a hash, verified byte-for-byte against the slot's representative in ActorVerify.
HD u64 fnv1a64(const u8* p, u32 len) { u64 h = 0xcbf29ce484222325ULL; for (u32 i = 0; i < len; i++) { h ^= p[i]; h *= 0x100000001b3ULL; } return h ? h : 1; }

struct ActorSlot { u64 hash; u64 first; /* (appRank << 32 | batch change), min wins */ u32 actorNum; u32 repOff; u32 repLen; u32 pad; };

HD u32 actor_find_or_insert(ActorSlot* slots, u64 mask, u64 h) {   // EMPTY32: the table is full (the host grows it and runs the kernel again)
  u64 s = mix64(h) & mask;
  for (u64 probes = 0; probes <= mask; probes++) {
    u64 cur = slots[s].hash;   // read first: after the first few inserts every lookup hits without an atomic
    if (cur == h) return (u32)s;
    if (cur == 0) { cur = atomic_cas(&slots[s].hash, (u64)0, h); if (cur == 0 || cur == h) return (u32)s; }
    s = (s + 1) & mask;
  }
  return EMPTY32;
}
HD u32 actor_find(const ActorSlot* slots, u64 mask, u64 h) {
  u64 s = mix64(h) & mask;
  for (u64 probes = 0; probes <= mask; probes++) { const u64 cur = slots[s].hash; if (cur == h) return (u32)s; if (cur == 0) return EMPTY32; s = (s + 1) & mask; }
  return EMPTY32;
}
// authors of applied changes claim slots; first (smallest application rank) appearance is recorded
struct ActorInternKernel {
  const u8* arena; const ChangeHot* meta; const u8* applied; const u32* appRank; ActorSlot* slots; u64 mask; u32* authorSlot; u32* full /* set when the table has no room: grown by the host, kernel run again */;
  HD void operator()(size_t b) const {
    if (!applied[b]) { authorSlot[b] = EMPTY32; return; }
    const u64 h = fnv1a64(arena + meta[b].actorOff, meta[b].actorLen);
    const u32 s = actor_find_or_insert(slots, mask, h);
    if (s == EMPTY32) { *full = 1; authorSlot[b] = EMPTY32; return; }
    const u64 cand = ((u64)appRank[b] << 32) | (u64)b;
    if (cand < slots[s].first) atomic_min(&slots[s].first, cand);
    authorSlot[b] = s;
  }
};
// resolves every (change, local actor index) to a slot; verifies bytes against the representative
struct ActorMapKernel {
  const u8* arena; const ChangeHot* meta; const u32* nActors; const u8* applied; const u32* appRank; const ActorSlot* slots; u64 mask;
  const u32* amapBase; u32* amap /* actorNum per (change, local index) */; u64* errWord;
  HD bool bytesEq(const u8* a, u32 la, const ActorSlot& s) const {
    if (la != s.repLen) return false;
    for (u32 i = 0; i < la; i++) if (a[i] != arena[s.repOff + i]) return false;
    return true;
  }
  HD void operator()(size_t b) const {
    if (!applied[b]) return;
    const u32 base = amapBase[b];
    ByteReader r(arena, meta[b].otherOff, meta[b].off + meta[b].len);
    for (u32 k = 0; k < nActors[b]; k++) {
      u32 off, len;
      if (k == 0) { off = meta[b].actorOff; len = meta[b].actorLen; }
      else { len = (u32)r.uleb(); off = r.pos; r.skip(len); }
      const u32 s = actor_find(slots, mask, fnv1a64(arena + off, len));
      // the actor must have authored a change that is applied no later than this one (new.js:1443-1447)
      if (s == EMPTY32 || slots[s].actorNum == EMPTY32 || (slots[s].first >> 32) > appRank[b]) { raise(errWord, KE_UNKNOWN_ACTOR, b); amap[base + k] = 0; continue; }
      if (!bytesEq(arena + off, len, slots[s])) { raise(errWord, KE_HASH_COLLISION, b); }
      amap[base + k] = slots[s].actorNum;
    }
  }
};

// ---------------------------------------------------------------- row finalisation (new.js:678-724 readNextChangeOp)
// Packs raw change-local columns into 64-bit ids with document actor numbers: id = ctr << 16 | actorNum.
struct OpRows {   // one entry per op of the batch
  u64 *id, *obj, *key;        // obj: 0 = _root; key: elemId (0 = _head) for list ops
  u32 *keyStrOff, *keyStrLen; // keyStrLen == NULL32: list op
  u32 *flags;                 // bit0 insert, bits 8..23 action (0xffff = null)
  u32 *valLen, *valOff, *predOff, *predNum, *change /* batch change index */, *time /* application time, 1-based */;
  u64 *predId;                // one entry per pred
};
HD u64 pack_id(u64 ctr, u32 actorNum) { return (ctr << 16) | (u64)actorNum; }
HD u64 id_ctr(u64 id) { return id >> 16; }
HD u32 id_actor(u64 id) { return (u32)(id & 0xffff); }
static const u32 F_INSERT = 1u;
HD u32 flags_action(u32 f) { return (f >> 8) & 0xffffu; }

struct FinalizeOpsKernel {
  size_t numChanges; const ChangeHot* meta; const u32* nActors; const u32* opBase /* first op of each applied change (masked scan) */;
  const u32* predBase; const u32* rawBase /* first raw row of each change (every change of the batch) */; const u32* rawPredBase;
  const u32* timeBase /* per change: first op's application time */;
  const u32* amapBase; const u32* amap; const u8* applied; RawRows raw; OpRows rows; u64* errWord;
  // binary search: change of op i
  HD size_t changeOf(u32 i) const {
    size_t lo = 0, hi = numChanges;   // largest c with opBase[c] <= i (changes without ops share their successor's base)
    while (hi - lo > 1) { size_t mid = (lo + hi) / 2; if (opBase[mid] <= i) lo = mid; else hi = mid; }
    return lo;
  }
  HD u32 actorOf(size_t c, u32 local, bool& bad) const {
    if (local >= nActors[c]) { bad = true; return 0; }
    return amap[amapBase[c] + local];
  }
  HD void operator()(size_t i) const {
    const size_t c = changeOf((u32)i);
    const u32 k = (u32)i - opBase[c]; bool bad = false;
    const u32 r = rawBase[c] + k;   // raw rows are in batch order for every change; ops in application-masked order
    const u32 author = amap[amapBase[c]];
    const u64 startOp = meta[c].startOp;
    rows.id[i] = pack_id(startOp + k, author);
    rows.change[i] = (u32)c; rows.time[i] = timeBase[c] + k;
    const u32 oa = raw.objActor[r], oc = raw.objCtr[r];
    if ((oc == NULL32) != (oa == NULL32)) raise(errWord, KE_OBJ_MISMATCH, c);
    rows.obj[i] = oc == NULL32 ? 0 : pack_id(oc, actorOf(c, oa == NULL32 ? 0 : oa, bad));
    const u32 ka = raw.keyActor[r], kc = raw.keyCtr[r];
    if ((kc == NULL32 && ka != NULL32) || (kc == 0 && ka != NULL32) || (kc != NULL32 && kc > 0 && ka == NULL32)) raise(errWord, KE_KEY_MISMATCH, c);
    rows.key[i] = (kc == NULL32 || kc == 0 || ka == NULL32) ? 0 : pack_id(kc, actorOf(c, ka, bad));
    rows.keyStrOff[i] = raw.keyStrOff[r]; rows.keyStrLen[i] = raw.keyStrLen[r];
    const u32 act = raw.action[r];
    rows.flags[i] = (raw.insert[r] ? F_INSERT : 0) | ((act == NULL32 ? 0xffffu : (act > 0xfffe ? 0xfffeu : act)) << 8);
    rows.valLen[i] = raw.valLen[r] == NULL32 ? 0 : raw.valLen[r]; rows.valOff[i] = raw.valOff[r];
    const u32 pn = raw.predNum[r], pr = raw.predOff[r], po = pr - rawPredBase[c] + predBase[c];
    rows.predOff[i] = po; rows.predNum[i] = pn;
    for (u32 j = 0; j < pn; j++) {
      const u32 pa = raw.predActor[pr + j], pc = raw.predCtr[pr + j];
      rows.predId[po + j] = (pa == NULL32 || pc == NULL32) ? 0 : pack_id(pc, actorOf(c, pa, bad));
    }
    if (bad) raise(errWord, KE_ACTOR_INDEX, c);
    if (startOp + k >= (1ULL << 47)) raise(errWord, KE_TOO_LARGE, c);
  }
};

}  // namespace amg
