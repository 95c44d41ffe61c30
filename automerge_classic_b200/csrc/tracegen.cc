// amgpu — synthetic edit-trace generator (bench / test input; host C++, no CUDA).
//
// Produces binary changes in automerge-classic's wire format (reference backend/columnar.js:710-739
// encodeChange; container :659-686; RLE / delta / boolean columns backend/encoding.js:558-783,
// 932-998, 1061-1135) for the workloads of SURVEY.md §8(d):
//   C1  100 x 1-op changes, `set _root.kNNN`, one actor
//   C2  makeText + n single-op insert changes, one actor (C2b: the same ops in ONE bulk change)
//   C3  makeText + A actors x n/A single-op changes, 70 % insert / 30 % delete, merge every 100 changes
//   C4  nested maps: A actors x rounds x 100-op `set` changes, Zipf keys, Lamport-conflict heavy
//   C7  counters in root keys: create / increment / overwrite / delete, concurrent writers
//   C8  C6 plus counter elements: inserted counters, concurrent increments, overwrites and deletes of them
//   C6  one list of scalars and map objects: inserts, element updates / conflicts / deletes, keys set inside element maps
// Seeded SplitMix64; actor k = first 16 bytes of SHA-256("amgpu-actor-" || seed || k).
// The oracle (tests) decodes and re-applies these bytes, which cross-checks this independent encoder.
#include <zlib.h>
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace {
typedef std::array<uint8_t, 32> Hash;

struct Sha256 {
  uint32_t h[8]; uint8_t block[64]; size_t fill = 0; uint64_t total = 0;
  Sha256() { static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19}; memcpy(h, iv, 32); }
  static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
  void compress(const uint8_t* p) {
    static const uint32_t K[64] = {
      0x428a2f98,0x71374491,0xb5c0fbcf,0xe9b5dba5,0x3956c25b,0x59f111f1,0x923f82a4,0xab1c5ed5,0xd807aa98,0x12835b01,0x243185be,0x550c7dc3,
      0x72be5d74,0x80deb1fe,0x9bdc06a7,0xc19bf174,0xe49b69c1,0xefbe4786,0x0fc19dc6,0x240ca1cc,0x2de92c6f,0x4a7484aa,0x5cb0a9dc,0x76f988da,
      0x983e5152,0xa831c66d,0xb00327c8,0xbf597fc7,0xc6e00bf3,0xd5a79147,0x06ca6351,0x14292967,0x27b70a85,0x2e1b2138,0x4d2c6dfc,0x53380d13,
      0x650a7354,0x766a0abb,0x81c2c92e,0x92722c85,0xa2bfe8a1,0xa81a664b,0xc24b8b70,0xc76c51a3,0xd192e819,0xd6990624,0xf40e3585,0x106aa070,
      0x19a4c116,0x1e376c08,0x2748774c,0x34b0bcb5,0x391c0cb3,0x4ed8aa4a,0x5b9cca4f,0x682e6ff3,0x748f82ee,0x78a5636f,0x84c87814,0x8cc70208,
      0x90befffa,0xa4506ceb,0xbef9a3f7,0xc67178f2};
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4*i] << 24 | (uint32_t)p[4*i+1] << 16 | (uint32_t)p[4*i+2] << 8 | p[4*i+3];
    for (int i = 16; i < 64; i++) { uint32_t s0 = rotr(w[i-15], 7) ^ rotr(w[i-15], 18) ^ (w[i-15] >> 3), s1 = rotr(w[i-2], 17) ^ rotr(w[i-2], 19) ^ (w[i-2] >> 10); w[i] = w[i-16] + s0 + w[i-7] + s1; }
    uint32_t a=h[0],b=h[1],c=h[2],d=h[3],e=h[4],f=h[5],g=h[6],hh=h[7];
    for (int i = 0; i < 64; i++) {
      uint32_t t1 = hh + (rotr(e,6)^rotr(e,11)^rotr(e,25)) + ((e&f)^(~e&g)) + K[i] + w[i], t2 = (rotr(a,2)^rotr(a,13)^rotr(a,22)) + ((a&b)^(a&c)^(b&c));
      hh=g; g=f; f=e; e=d+t1; d=c; c=b; b=a; a=t1+t2;
    }
    h[0]+=a; h[1]+=b; h[2]+=c; h[3]+=d; h[4]+=e; h[5]+=f; h[6]+=g; h[7]+=hh;
  }
  void update(const uint8_t* p, size_t n) { total += n; while (n) { size_t t = std::min(n, 64 - fill); memcpy(block + fill, p, t); fill += t; p += t; n -= t; if (fill == 64) { compress(block); fill = 0; } } }
  Hash digest() { uint64_t bits = total * 8; uint8_t pad = 0x80; update(&pad, 1); uint8_t z = 0; while (fill != 56) update(&z, 1); uint8_t l[8]; for (int i = 0; i < 8; i++) l[i] = (uint8_t)(bits >> (56 - 8*i)); update(l, 8);
    Hash o; for (int i = 0; i < 8; i++) { o[4*i] = h[i] >> 24; o[4*i+1] = h[i] >> 16; o[4*i+2] = h[i] >> 8; o[4*i+3] = h[i]; } return o; }
};

struct Rng { uint64_t s; uint64_t next() { uint64_t z = (s += 0x9e3779b97f4a7c15ULL); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }
  uint64_t below(uint64_t n) { return next() % n; } double unit() { return (next() >> 11) * (1.0 / 9007199254740992.0); } };

typedef std::string Bytes;
void uleb(Bytes& b, uint64_t v) { do { uint8_t x = v & 0x7f; v >>= 7; if (v) x |= 0x80; b.push_back((char)x); } while (v); }
void sleb(Bytes& b, int64_t v) { while (true) { uint8_t x = v & 0x7f; v >>= 7; if ((v == 0 && !(x & 0x40)) || (v == -1 && (x & 0x40))) { b.push_back((char)x); return; } b.push_back((char)(x | 0x80)); } }

static const int64_t NUL = INT64_MIN;
// canonical RLE of int64 values (NUL = null); `isSigned` selects sLEB vs uLEB raw values
Bytes rle(const std::vector<int64_t>& v, bool isSigned) {
  Bytes out; bool allNull = true; for (auto x : v) if (x != NUL) allNull = false;
  if (allNull) return out;
  std::vector<int64_t> lit;
  auto raw = [&](int64_t x) { if (isSigned) sleb(out, x); else uleb(out, (uint64_t)x); };
  auto flush = [&]() { if (lit.empty()) return; sleb(out, -(int64_t)lit.size()); for (auto x : lit) raw(x); lit.clear(); };
  for (size_t i = 0; i < v.size();) {
    size_t j = i; while (j < v.size() && v[j] == v[i]) j++;
    const int64_t n = (int64_t)(j - i);
    if (v[i] == NUL) { flush(); sleb(out, 0); uleb(out, (uint64_t)n); }
    else if (n >= 2) { flush(); sleb(out, n); raw(v[i]); }
    else lit.push_back(v[i]);
    i = j;
  }
  flush(); return out;
}
Bytes delta(const std::vector<int64_t>& v) { std::vector<int64_t> d; int64_t last = 0; for (auto x : v) { if (x == NUL) d.push_back(NUL); else { d.push_back(x - last); last = x; } } return rle(d, true); }
Bytes rleStr(const std::vector<const std::string*>& v) {   // nullptr = null
  Bytes out; bool allNull = true; for (auto x : v) if (x) allNull = false;
  if (allNull) return out;
  std::vector<const std::string*> lit;
  auto raw = [&](const std::string* s) { uleb(out, s->size()); out += *s; };
  auto flush = [&]() { if (lit.empty()) return; sleb(out, -(int64_t)lit.size()); for (auto x : lit) raw(x); lit.clear(); };
  auto eq = [](const std::string* a, const std::string* b) { return (!a && !b) || (a && b && *a == *b); };
  for (size_t i = 0; i < v.size();) {
    size_t j = i; while (j < v.size() && eq(v[j], v[i])) j++;
    const int64_t n = (int64_t)(j - i);
    if (!v[i]) { flush(); sleb(out, 0); uleb(out, (uint64_t)n); } else if (n >= 2) { flush(); sleb(out, n); raw(v[i]); } else lit.push_back(v[i]);
    i = j;
  }
  flush(); return out;
}
Bytes boolean(const std::vector<uint8_t>& v) { Bytes out; bool last = false; uint64_t cnt = 0; for (auto x : v) { if ((x != 0) == last) cnt++; else { uleb(out, cnt); last = x != 0; cnt = 1; } } if (cnt > 0) uleb(out, cnt); return out; }

struct OpId { uint64_t ctr = 0; int actor = -1; };   // actor = global actor index
struct Op {
  int action; OpId obj;   // obj.actor < 0: root
  bool isMapKey = false; std::string key; OpId elem; bool insert = false;   // elem.ctr == 0: _head
  bool hasValue = false; int valTag = 0; Bytes valRaw;
  std::vector<OpId> pred;
};
struct Actor { Bytes id; };

struct Encoder {
  const std::vector<Actor>* actors;
  Bytes encode(int author, uint64_t seq, uint64_t startOp, const std::vector<Hash>& depsIn, const std::vector<Op>& ops, Hash* hashOut, bool deflate = true) const {
    // change-local actor table: author first, then the other referenced actors sorted by id (columnar.js:154-157)
    std::vector<int> others;
    auto note = [&](int a) { if (a >= 0 && a != author && std::find(others.begin(), others.end(), a) == others.end()) others.push_back(a); };
    for (auto& op : ops) { note(op.obj.actor); if (!op.isMapKey) note(op.elem.ctr ? op.elem.actor : -1); for (auto& p : op.pred) note(p.actor); }
    std::sort(others.begin(), others.end(), [&](int a, int b) { return (*actors)[a].id < (*actors)[b].id; });
    auto local = [&](int a) -> int64_t { if (a == author) return 0; return 1 + (std::find(others.begin(), others.end(), a) - others.begin()); };
    std::vector<int64_t> objActor, objCtr, keyActor, keyCtr, action, valLen, predNum, predActor, predCtr, chld; std::vector<const std::string*> keyStr; std::vector<uint8_t> ins; Bytes valRaw;
    for (auto& op : ops) {
      if (op.obj.actor < 0) { objActor.push_back(NUL); objCtr.push_back(NUL); } else { objActor.push_back(local(op.obj.actor)); objCtr.push_back((int64_t)op.obj.ctr); }
      if (op.isMapKey) { keyActor.push_back(NUL); keyCtr.push_back(NUL); keyStr.push_back(&op.key); }
      else if (op.elem.ctr == 0) { keyActor.push_back(NUL); keyCtr.push_back(0); keyStr.push_back(nullptr); }
      else { keyActor.push_back(local(op.elem.actor)); keyCtr.push_back((int64_t)op.elem.ctr); keyStr.push_back(nullptr); }
      ins.push_back(op.insert); action.push_back(op.action);
      if (op.hasValue) { valLen.push_back((int64_t)(op.valRaw.size() << 4 | (unsigned)op.valTag)); valRaw += op.valRaw; } else valLen.push_back(0);
      chld.push_back(NUL);
      std::vector<OpId> preds = op.pred;
      std::sort(preds.begin(), preds.end(), [&](const OpId& a, const OpId& b) { return a.ctr != b.ctr ? a.ctr < b.ctr : (*actors)[a.actor].id < (*actors)[b.actor].id; });
      predNum.push_back((int64_t)preds.size()); for (auto& p : preds) { predActor.push_back(local(p.actor)); predCtr.push_back((int64_t)p.ctr); }
    }
    std::vector<std::pair<int, Bytes>> cols = {
      {0x01, rle(objActor, false)}, {0x02, rle(objCtr, false)}, {0x11, rle(keyActor, false)}, {0x13, delta(keyCtr)}, {0x15, rleStr(keyStr)},
      {0x34, boolean(ins)}, {0x42, rle(action, false)}, {0x56, rle(valLen, false)}, {0x57, valRaw}, {0x61, rle(chld, false)}, {0x63, delta(chld)},
      {0x70, rle(predNum, false)}, {0x71, rle(predActor, false)}, {0x73, delta(predCtr)}};
    Bytes body; std::vector<Hash> deps = depsIn; std::sort(deps.begin(), deps.end());
    uleb(body, deps.size()); for (auto& d : deps) body.append((const char*)d.data(), 32);
    const Bytes& aid = (*actors)[author].id; uleb(body, aid.size()); body += aid;
    uleb(body, seq); uleb(body, startOp); sleb(body, 0); uleb(body, 0);
    uleb(body, others.size()); for (int o : others) { uleb(body, (*actors)[o].id.size()); body += (*actors)[o].id; }
    size_t nonEmpty = 0; for (auto& c : cols) if (!c.second.empty()) nonEmpty++;
    uleb(body, nonEmpty); for (auto& c : cols) if (!c.second.empty()) { uleb(body, (uint64_t)c.first); uleb(body, c.second.size()); }
    for (auto& c : cols) body += c.second;
    Bytes header; header.push_back(1); uleb(header, body.size());
    Sha256 s; s.update((const uint8_t*)header.data(), header.size()); s.update((const uint8_t*)body.data(), body.size()); Hash h = s.digest(); if (hashOut) *hashOut = h;
    static const uint8_t magic[4] = {0x85, 0x6f, 0x4a, 0x83};
    Bytes out((const char*)magic, 4); out.append((const char*)h.data(), 4);
    if (deflate && 8 + header.size() + body.size() >= 256) {   // columnar.js:738, 798-808 (zlib level 6 raw deflate; bytes need not match pako's)
      z_stream zs; memset(&zs, 0, sizeof(zs)); deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
      Bytes comp; comp.resize(deflateBound(&zs, body.size()));
      zs.next_in = (Bytef*)body.data(); zs.avail_in = (uInt)body.size(); zs.next_out = (Bytef*)comp.data(); zs.avail_out = (uInt)comp.size();
      ::deflate(&zs, Z_FINISH); comp.resize(zs.total_out); deflateEnd(&zs);
      out.push_back(2); uleb(out, comp.size()); out += comp;
    } else { out += header; out += body; }
    return out;
  }
};

struct Trace { Bytes blob; std::vector<uint64_t> offsets{0}; uint64_t totalOps = 0;
  void add(const Bytes& c, uint64_t ops) { blob += c; offsets.push_back(blob.size()); totalOps += ops; } };

std::vector<Actor> makeActors(uint64_t seed, int n) {
  std::vector<Actor> a(n);
  for (int k = 0; k < n; k++) { std::string s = "amgpu-actor-" + std::to_string(seed) + "-" + std::to_string(k); Sha256 h; h.update((const uint8_t*)s.data(), s.size()); Hash d = h.digest(); a[k].id.assign((const char*)d.data(), 16); }
  return a;
}
Op charInsert(OpId text, OpId after, char c) { Op op; op.action = 1; op.obj = text; op.elem = after; op.insert = true; op.hasValue = true; op.valTag = 6; op.valRaw.assign(1, c); return op; }

// C2 / C2b / C3: text trace
void genText(Trace& t, uint64_t seed, uint64_t nOps, int A, bool bulk, double delFrac, int perRound) {
  Rng rng{seed}; std::vector<Actor> actors = makeActors(seed, A); Encoder enc{&actors};
  std::vector<uint64_t> seq(A, 0); std::vector<Hash> lastHash(A); std::vector<bool> hasHash(A, false);
  OpId text{1, 0};
  { Op mk; mk.action = 4; mk.isMapKey = true; mk.key = "text"; Hash h; t.add(enc.encode(0, ++seq[0], 1, {}, {mk}, &h), 1); lastHash[0] = h; hasHash[0] = true; }
  uint64_t maxOp = 1;
  std::vector<OpId> visible; std::unordered_map<uint64_t, size_t> slotOf;   // key = ctr * 65536 + actor
  auto keyOf = [](const OpId& o) { return o.ctr * 65536 + (uint64_t)o.actor; };
  const uint64_t perActor = nOps / A;
  if (bulk) {
    std::vector<Op> ops; OpId last{0, 0}; std::vector<OpId> all;
    for (uint64_t i = 0; i < nOps; i++) {
      const double u = rng.unit(); OpId after = last;
      if (u >= 0.90 && u < 0.99 && !all.empty()) after = all[rng.below(all.size())]; else if (u >= 0.99) after = OpId{0, 0};
      ops.push_back(charInsert(text, after, (char)('a' + rng.below(26)))); last = OpId{maxOp + 1 + i, 0}; all.push_back(last);
    }
    Hash h; t.add(enc.encode(0, ++seq[0], maxOp + 1, {lastHash[0]}, ops, &h), nOps); return;
  }
  std::vector<OpId> lastIns(A, OpId{0, 0});
  uint64_t done = 0;
  for (uint64_t round = 0; done < perActor; round++) {
    const uint64_t inRound = std::min<uint64_t>(perRound, perActor - done);
    std::vector<Hash> roundHeads; for (int a = 0; a < A; a++) if (hasHash[a]) roundHeads.push_back(lastHash[a]);
    std::vector<std::vector<OpId>> ins(A), del(A); const uint64_t base = maxOp; uint64_t newMax = maxOp;
    for (int a = 0; a < A; a++) {
      std::vector<OpId>& myIns = ins[a]; std::vector<OpId>& myDel = del[a]; std::unordered_map<uint64_t, bool> deletedLocal; size_t aliveLocal = 0;
      for (uint64_t j = 0; j < inRound; j++) {
        const uint64_t ctr = base + 1 + j; Op op; const size_t viewSize = visible.size() - 0 + aliveLocal;
        const bool doDel = rng.unit() < delFrac && viewSize > deletedLocal.size() + 1;
        if (doDel) {
          OpId target; int tries = 0;
          while (true) {
            const uint64_t r = rng.below(visible.size() + myIns.size());
            target = r < visible.size() ? visible[r] : myIns[r - visible.size()];
            if (!deletedLocal.count(keyOf(target)) || ++tries > 64) break;
          }
          if (deletedLocal.count(keyOf(target))) { op = charInsert(text, lastIns[a], (char)('a' + rng.below(26))); lastIns[a] = OpId{ctr, a}; myIns.push_back(lastIns[a]); aliveLocal++; }
          else { op.action = 3; op.obj = text; op.elem = target; op.pred = {target}; deletedLocal[keyOf(target)] = true; myDel.push_back(target); }
        } else {
          const double u = rng.unit(); OpId after = lastIns[a];
          if (u >= 0.90 && u < 0.99 && visible.size() + myIns.size() > 0) { const uint64_t r = rng.below(visible.size() + myIns.size()); after = r < visible.size() ? visible[r] : myIns[r - visible.size()]; }
          else if (u >= 0.99) after = OpId{0, 0};
          op = charInsert(text, after, (char)('a' + rng.below(26))); lastIns[a] = OpId{ctr, a}; myIns.push_back(lastIns[a]); aliveLocal++;
        }
        std::vector<Hash> deps; if (j == 0) deps = roundHeads; else deps = {lastHash[a]};
        if (j == 0 && hasHash[a] && std::find(deps.begin(), deps.end(), lastHash[a]) == deps.end()) deps.push_back(lastHash[a]);
        Hash h; t.add(enc.encode(a, ++seq[a], ctr, deps, {op}, &h), 1); lastHash[a] = h; hasHash[a] = true; newMax = std::max(newMax, ctr);
      }
    }
    // merge: every actor now sees every insert and delete of the round
    for (int a = 0; a < A; a++) for (auto& e : ins[a]) { slotOf[keyOf(e)] = visible.size(); visible.push_back(e); }
    for (int a = 0; a < A; a++) for (auto& e : del[a]) {
      auto it = slotOf.find(keyOf(e)); if (it == slotOf.end()) continue;
      const size_t s = it->second; slotOf.erase(it);
      if (s != visible.size() - 1) { visible[s] = visible.back(); slotOf[keyOf(visible[s])] = s; }
      visible.pop_back();
    }
    maxOp = newMax; done += inRound;
  }
}

void genC1(Trace& t, uint64_t seed) {
  std::vector<Actor> actors = makeActors(seed, 1); Encoder enc{&actors}; Hash last; bool has = false;
  for (int i = 0; i < 100; i++) {
    Op op; op.action = 1; op.isMapKey = true; char k[8]; snprintf(k, sizeof k, "k%03d", i); op.key = k; op.hasValue = true; op.valTag = 4; sleb(op.valRaw, i);
    Hash h; t.add(enc.encode(0, i + 1, i + 1, has ? std::vector<Hash>{last} : std::vector<Hash>{}, {op}, &h), 1); last = h; has = true;
  }
}

// C4: nested maps, Zipf keys, conflicts between same-round writers
void genC4(Trace& t, uint64_t seed, uint64_t nOps, int A, int nChild, int nKeysPerChild, int opsPerChange) {
  Rng rng{seed}; std::vector<Actor> actors = makeActors(seed, A); Encoder enc{&actors};
  std::vector<uint64_t> seq(A, 0); std::vector<Hash> lastHash(A); std::vector<bool> hasHash(A, false);
  std::vector<Op> setup; { Op m; m.action = 0; m.isMapKey = true; m.key = "m"; setup.push_back(m); }
  for (int c = 0; c < nChild; c++) { Op m; m.action = 0; m.obj = OpId{1, 0}; m.isMapKey = true; char k[8]; snprintf(k, sizeof k, "c%02d", c); m.key = k; setup.push_back(m); }
  { Hash h; t.add(enc.encode(0, ++seq[0], 1, {}, setup, &h), setup.size()); lastHash[0] = h; hasHash[0] = true; }
  uint64_t maxOp = setup.size();
  const size_t K = (size_t)nChild * nKeysPerChild;
  std::vector<double> cdf(K); { double s = 0; for (size_t i = 0; i < K; i++) { s += 1.0 / (double)(i + 1); cdf[i] = s; } for (auto& x : cdf) x /= s; }
  std::vector<std::vector<OpId>> visible(K);
  const uint64_t rounds = std::max<uint64_t>(1, nOps / ((uint64_t)A * opsPerChange));
  for (uint64_t r = 0; r < rounds; r++) {
    std::vector<Hash> roundHeads; for (int a = 0; a < A; a++) if (hasHash[a]) roundHeads.push_back(lastHash[a]);
    std::unordered_map<size_t, std::vector<OpId>> written; const uint64_t base = maxOp;
    for (int a = 0; a < A; a++) {
      std::unordered_map<size_t, OpId> mine; std::vector<Op> ops;
      for (int j = 0; j < opsPerChange; j++) {
        const size_t key = std::lower_bound(cdf.begin(), cdf.end(), rng.unit()) - cdf.begin(); const size_t kk = std::min(key, K - 1);
        Op op; op.action = 1; op.obj = OpId{2 + kk / nKeysPerChild, 0}; op.isMapKey = true; char k[8]; snprintf(k, sizeof k, "k%02d", (int)(kk % nKeysPerChild)); op.key = k;
        op.hasValue = true; op.valTag = 4; sleb(op.valRaw, (int64_t)rng.below(1000000));
        auto it = mine.find(kk); if (it != mine.end()) op.pred = {it->second}; else op.pred = visible[kk];
        mine[kk] = OpId{base + 1 + j, a}; ops.push_back(op);
      }
      std::vector<Hash> deps = roundHeads; if (hasHash[a] && std::find(deps.begin(), deps.end(), lastHash[a]) == deps.end()) deps.push_back(lastHash[a]);
      Hash h; t.add(enc.encode(a, ++seq[a], base + 1, deps, ops, &h), ops.size()); lastHash[a] = h; hasHash[a] = true;
      for (auto& kv : mine) written[kv.first].push_back(kv.second);
    }
    for (auto& kv : written) visible[kv.first] = kv.second;
    maxOp = base + opsPerChange;
  }
}

// C6: one list of mixed elements: scalar inserts, inserted map objects with keys set inside them, value updates and
// concurrent conflicting updates of existing elements, conflict-adding sets (empty pred), deletes, several ops of one
// change on the same element. Every actor works on the state of the round start (so same-round writers conflict).
void genRichList(Trace& t, uint64_t seed, uint64_t nOps, int A, int maxOpsPerChange, bool withCounters = false) {
  Rng rng{seed}; std::vector<Actor> actors = makeActors(seed, A); Encoder enc{&actors};
  std::vector<uint64_t> seq(A, 0); std::vector<Hash> lastHash(A); std::vector<bool> hasHash(A, false);
  OpId list{1, 0};
  { Op mk; mk.action = 2; mk.isMapKey = true; mk.key = "items"; Hash h; t.add(enc.encode(0, ++seq[0], 1, {}, {mk}, &h), 1); lastHash[0] = h; hasHash[0] = true; }
  uint64_t maxOp = 1, produced = 0;
  struct Row { OpId id; bool isMap; bool isCounter = false; };
  struct Elem { OpId id; std::vector<Row> vis; };
  std::vector<Elem> elems; std::unordered_map<uint64_t, size_t> elemAt;
  std::unordered_map<uint64_t, std::vector<OpId>> mapKeyVis;   // (map object key * 8 + key index) -> visible set rows
  auto keyOf = [](const OpId& o) { return o.ctr * 65536 + (uint64_t)o.actor; };
  auto intValue = [&](Op& op) { op.hasValue = true; op.valTag = 4; sleb(op.valRaw, (int64_t)rng.below(100000)); };
  while (produced < nOps) {
    std::vector<Hash> roundHeads; for (int a = 0; a < A; a++) if (hasHash[a]) roundHeads.push_back(lastHash[a]);
    const uint64_t base = maxOp; uint64_t newMax = maxOp;
    struct Pending { std::vector<Elem> newElems; std::vector<std::pair<size_t, Row>> newRows; std::vector<OpId> overwritten; std::vector<std::pair<uint64_t, OpId>> mapSets; };
    std::vector<Pending> pend(A);
    for (int a = 0; a < A && produced < nOps; a++) {
      Pending& P = pend[a]; std::vector<Op> ops; const int want = 1 + (int)rng.below(maxOpsPerChange);
      std::unordered_map<size_t, std::vector<Row>> localVis;     // element slot -> rows visible to this change so far
      std::unordered_map<uint64_t, std::vector<OpId>> localKey;  // nested map key -> rows visible to this change so far
      auto visOf = [&](size_t e) -> std::vector<Row>& { auto it = localVis.find(e); if (it == localVis.end()) it = localVis.emplace(e, elems[e].vis).first; return it->second; };
      OpId lastIns{0, 0}; long lastTouched = -1;
      while ((int)ops.size() < want) {
        const uint64_t ctr = base + 1 + ops.size(); const OpId me{ctr, a}; const double u = rng.unit(); Op op; op.obj = list;
        size_t e = elems.empty() ? 0 : (size_t)rng.below(elems.size());
        if (lastTouched >= 0 && rng.unit() < 0.25) e = (size_t)lastTouched;   // several ops of one change on the same element
        if (withCounters && !elems.empty() && rng.unit() < 0.22) {   // counters (C8): insert one, or increment a visible one
          std::vector<Row>& v = visOf(e); const Row* cnt = nullptr; for (auto& r : v) if (r.isCounter) cnt = &r;
          if (cnt && rng.unit() < 0.7) {
            op.action = 5; op.elem = elems[e].id; op.hasValue = true; const bool neg = rng.unit() < 0.3; op.valTag = neg ? 4 : 3;
            if (neg) sleb(op.valRaw, -(int64_t)rng.below(20)); else uleb(op.valRaw, rng.below(100));
            op.pred = {cnt->id}; lastTouched = (long)e; ops.push_back(op);
          } else {
            op.action = 1; op.insert = true; op.hasValue = true; op.valTag = 8; sleb(op.valRaw, (int64_t)rng.below(50));
            op.elem = rng.unit() < 0.5 ? lastIns : elems[e].id;
            Row nr{me, false}; nr.isCounter = true; P.newElems.push_back(Elem{me, {nr}}); lastIns = me; ops.push_back(op);
          }
          continue;
        }
        if (elems.empty() || u < 0.35) {            // insert a scalar
          op.action = 1; op.insert = true; intValue(op);
          const double w = rng.unit(); op.elem = w < 0.6 ? lastIns : (w < 0.95 && !elems.empty() ? elems[rng.below(elems.size())].id : OpId{0, 0});
          P.newElems.push_back(Elem{me, {Row{me, false}}}); lastIns = me; ops.push_back(op);
        } else if (u < 0.45) {                      // insert a map object and set a key inside it
          op.action = 0; op.insert = true; op.elem = rng.unit() < 0.5 ? lastIns : elems[e].id;
          P.newElems.push_back(Elem{me, {Row{me, true}}}); lastIns = me; ops.push_back(op);
          if ((int)ops.size() < want) { Op st; st.action = 1; st.obj = me; st.isMapKey = true; st.key = "k0"; intValue(st); const OpId sid{ctr + 1, a}; P.mapSets.push_back({keyOf(me) * 8, sid}); localKey[keyOf(me) * 8] = {sid}; ops.push_back(st); }
        } else if (u < 0.70) {                      // overwrite the value(s) of an element (also re-inserts a concurrently deleted one)
          std::vector<Row>& v = visOf(e); op.action = rng.unit() < 0.15 ? 0 : 1; op.elem = elems[e].id; if (op.action == 1) intValue(op);
          for (auto& r : v) { op.pred.push_back(r.id); P.overwritten.push_back(r.id); }
          v.clear(); v.push_back(Row{me, op.action == 0}); P.newRows.push_back({e, v.back()}); lastTouched = (long)e; ops.push_back(op);
        } else if (u < 0.75) {                      // add a conflicting value without overwriting anything
          std::vector<Row>& v = visOf(e); op.action = 1; op.elem = elems[e].id; intValue(op);
          v.push_back(Row{me, false}); P.newRows.push_back({e, v.back()}); lastTouched = (long)e; ops.push_back(op);
        } else if (u < 0.88) {                      // delete an element
          std::vector<Row>& v = visOf(e); if (v.empty()) continue;
          op.action = 3; op.elem = elems[e].id; for (auto& r : v) { op.pred.push_back(r.id); P.overwritten.push_back(r.id); }
          v.clear(); lastTouched = (long)e; ops.push_back(op);
        } else {                                    // set a key inside a map that lives in a list element
          std::vector<Row>& v = visOf(e); const Row* m = nullptr; for (auto& r : v) if (r.isMap) m = &r;
          if (!m) continue;
          const uint64_t kk = keyOf(m->id) * 8 + rng.below(3); op.obj = m->id; op.action = 1; op.isMapKey = true; op.key = std::string("k") + (char)('0' + kk % 8); intValue(op);
          auto it = localKey.find(kk); if (it == localKey.end()) it = localKey.emplace(kk, mapKeyVis[kk]).first;
          for (auto& r : it->second) { op.pred.push_back(r); P.overwritten.push_back(r); }
          it->second = {me}; P.mapSets.push_back({kk, me}); ops.push_back(op);
        }
      }
      std::vector<Hash> deps = roundHeads; if (hasHash[a] && std::find(deps.begin(), deps.end(), lastHash[a]) == deps.end()) deps.push_back(lastHash[a]);
      Hash h; t.add(enc.encode(a, ++seq[a], base + 1, deps, ops, &h), ops.size()); lastHash[a] = h; hasHash[a] = true;
      produced += ops.size(); newMax = std::max<uint64_t>(newMax, base + ops.size());
    }
    // merge the round: rows named in any pred disappear, everything else written this round becomes visible
    std::unordered_map<uint64_t, bool> gone; for (auto& P : pend) for (auto& o : P.overwritten) gone[keyOf(o)] = true;
    for (auto& el : elems) { std::vector<Row> keep; for (auto& r : el.vis) if (!gone.count(keyOf(r.id))) keep.push_back(r); el.vis.swap(keep); }
    for (auto& kv : mapKeyVis) { std::vector<OpId> keep; for (auto& r : kv.second) if (!gone.count(keyOf(r))) keep.push_back(r); kv.second.swap(keep); }
    for (auto& P : pend) {
      for (auto& nr : P.newRows) if (!gone.count(keyOf(nr.second.id))) elems[nr.first].vis.push_back(nr.second);
      for (auto& ms : P.mapSets) if (!gone.count(keyOf(ms.second))) mapKeyVis[ms.first].push_back(ms.second);
      for (auto& ne : P.newElems) { Elem el = ne; if (gone.count(keyOf(el.id))) el.vis.clear(); elemAt[keyOf(el.id)] = elems.size(); elems.push_back(el); }
    }
    maxOp = newMax;
  }
}

// C7: counters in maps: `set` of datatype counter, concurrent `inc` ops (pred = the counter's set op), deletes and plain
// overwrites of counter keys; a few root keys so that same-round writers collide.
void genCounters(Trace& t, uint64_t seed, uint64_t nOps, int A, int nKeys, int maxOpsPerChange) {
  Rng rng{seed}; std::vector<Actor> actors = makeActors(seed, A); Encoder enc{&actors};
  std::vector<uint64_t> seq(A, 0); std::vector<Hash> lastHash(A); std::vector<bool> hasHash(A, false);
  struct Val { OpId id; bool counter; };
  std::vector<std::vector<Val>> visible(nKeys);     // visible set ops per key (an `inc` never becomes a value of its own)
  uint64_t maxOp = 0, produced = 0;
  auto keyOf = [](const OpId& o) { return o.ctr * 65536 + (uint64_t)o.actor; };
  while (produced < nOps) {
    std::vector<Hash> roundHeads; for (int a = 0; a < A; a++) if (hasHash[a]) roundHeads.push_back(lastHash[a]);
    const uint64_t base = maxOp; uint64_t newMax = maxOp;
    std::vector<std::vector<std::pair<int, Val>>> written(A); std::vector<OpId> overwritten;
    for (int a = 0; a < A && produced < nOps; a++) {
      std::vector<Op> ops; const int want = 1 + (int)rng.below(maxOpsPerChange);
      std::unordered_map<int, std::vector<Val>> local;
      auto visOf = [&](int k) -> std::vector<Val>& { auto it = local.find(k); if (it == local.end()) it = local.emplace(k, visible[k]).first; return it->second; };
      while ((int)ops.size() < want) {
        const uint64_t ctr = base + 1 + ops.size(); const OpId me{ctr, a}; const int k = (int)rng.below(nKeys); std::vector<Val>& v = visOf(k);
        Op op; op.isMapKey = true; char kb[8]; snprintf(kb, sizeof kb, "c%02d", k); op.key = kb; const double u = rng.unit();
        const Val* counter = nullptr; for (auto& x : v) if (x.counter) counter = &x;
        if (counter && u < 0.55) {            // increment (possibly negative)
          op.action = 5; op.hasValue = true; const bool neg = rng.unit() < 0.3; op.valTag = neg ? 4 : 3;
          if (neg) sleb(op.valRaw, -(int64_t)rng.below(50)); else uleb(op.valRaw, rng.below(1000));
          op.pred = {counter->id};
        } else if (u < 0.80) {                // (re)create a counter, overwriting whatever is visible
          op.action = 1; op.hasValue = true; op.valTag = 8; sleb(op.valRaw, (int64_t)rng.below(100));
          for (auto& x : v) { op.pred.push_back(x.id); overwritten.push_back(x.id); }
          v.clear(); v.push_back(Val{me, true}); written[a].push_back({k, v.back()});
        } else if (u < 0.90) {                // plain value
          op.action = 1; op.hasValue = true; op.valTag = 4; sleb(op.valRaw, (int64_t)rng.below(100000));
          for (auto& x : v) { op.pred.push_back(x.id); overwritten.push_back(x.id); }
          v.clear(); v.push_back(Val{me, false}); written[a].push_back({k, v.back()});
        } else {                              // delete the key
          if (v.empty()) continue;
          op.action = 3; for (auto& x : v) { op.pred.push_back(x.id); overwritten.push_back(x.id); }
          v.clear();
        }
        ops.push_back(op);
      }
      std::vector<Hash> deps = roundHeads; if (hasHash[a] && std::find(deps.begin(), deps.end(), lastHash[a]) == deps.end()) deps.push_back(lastHash[a]);
      Hash h; t.add(enc.encode(a, ++seq[a], base + 1, deps, ops, &h), ops.size()); lastHash[a] = h; hasHash[a] = true;
      produced += ops.size(); newMax = std::max<uint64_t>(newMax, base + ops.size());
    }
    std::unordered_map<uint64_t, bool> gone; for (auto& o : overwritten) gone[keyOf(o)] = true;
    for (auto& vs : visible) { std::vector<Val> keep; for (auto& x : vs) if (!gone.count(keyOf(x.id))) keep.push_back(x); vs.swap(keep); }
    for (auto& w : written) for (auto& kv : w) if (!gone.count(keyOf(kv.second.id))) visible[kv.first].push_back(kv.second);
    maxOp = newMax;
  }
}
}  // namespace

extern "C" {
// config: 1 = C1, 2 = C2, 22 = C2b (bulk), 3 = C3, 4 = C4, 6 = C6 (rich list), 7 = C7 (counters), 8 = C8 (rich list with counter elements). Returns malloc'ed blob + offsets (n_changes + 1).
int amg_trace_generate(int config, uint64_t seed, uint64_t n_ops, int n_actors, uint8_t** blob, size_t* blob_len, uint64_t** offsets, size_t* n_changes, uint64_t* total_ops) {
  Trace t;
  if (config == 1) genC1(t, seed);
  else if (config == 2) genText(t, seed, n_ops, 1, false, 0.0, 100);
  else if (config == 22) genText(t, seed, n_ops, 1, true, 0.0, 100);
  else if (config == 3) genText(t, seed, n_ops, n_actors > 0 ? n_actors : 10, false, 0.3, 100);
  else if (config == 4) genC4(t, seed, n_ops, n_actors > 0 ? n_actors : 100, 100, 100, 100);
  else if (config == 7) genCounters(t, seed, n_ops, n_actors > 0 ? n_actors : 3, 6, 5);
  else if (config == 8) genRichList(t, seed, n_ops, n_actors > 0 ? n_actors : 3, 6, true);
  else if (config == 6) genRichList(t, seed, n_ops, n_actors > 0 ? n_actors : 4, 6);
  else return 1;
  *blob = (uint8_t*)malloc(t.blob.size() + 64); memcpy(*blob, t.blob.data(), t.blob.size()); memset(*blob + t.blob.size(), 0, 64); *blob_len = t.blob.size();
  *offsets = (uint64_t*)malloc(t.offsets.size() * 8); memcpy(*offsets, t.offsets.data(), t.offsets.size() * 8);
  *n_changes = t.offsets.size() - 1; *total_ops = t.totalOps; return 0;
}
void amg_trace_free(void* p) { free(p); }
}
