// amgpu — kernels #4: patch emission.
//
// Replaces (reference paths relative to /root/reference):
//   backend/new.js:884-1040   updatePatchProperty   backend/new.js:747-782  appendEdit (run coalescing)
//   backend/new.js:1461-1528  setupPatches          backend/new.js:1604-1635 documentPatch (getPatch)
//
// The reference's incremental patch is an edit log in application order whose list indexes are valid
// at the moment each edit is appended. For an op applied at time t on the element at document
// position p the index is  #{elements e of the same list : pos_e < p, t_ins(e) < t, t_del(e) > t}.
// That is an offline dominance count; it is computed for all ops at once by a most-significant-bit
// first radix partition over the time bits (DominanceLevelKernel): items (element insertions +1,
// first deletions -1, and one query per op) start in (object, position) order; at bit b every query
// with bit b set adds the weighted count of bit-b-clear points that precede it inside its current
// group (same higher time bits, same list), then the group is stably split on bit b.
// Supported in incremental mode: map/table objects (set/del/make*, conflicts) and list/text objects
// whose touched elements carry only their insert row (insert + delete; no element updates, no nested
// objects inside lists, no counters). Anything else raises AMG_ERR_UNSUPPORTED — there is no CPU path.
#pragma once
#include "opset.cuh"

namespace amg {

// flat patch records (copied to the host verbatim)
struct PropRec { u64 obj, opId; u32 keyOff, keyLen, valLen, valOff, flags /* action<<8 | 1 = empty key */, pad; };
struct EditRec { u64 obj, opId; u32 index, kind /* 0 insert 1 remove 2 update | runStart<<8 | action<<16 */, valLen, valOff; };
enum { EK_INSERT = 0, EK_REMOVE = 1, EK_UPDATE = 2 };

// ---------------------------------------------------------------- per-position state in document order
struct GroupHeadKernel {   // group = rows of one map key / one list element (insert row + its update rows), adjacent in document order
  const u8* arena; DocRows d; u32* head;
  HD void operator()(size_t p) const {
    bool h = p == 0 || d.obj[p] != d.obj[p - 1];
    if (!h) {
      const bool listA = d.keyStrLen[p] == NULL32, listB = d.keyStrLen[p - 1] == NULL32;
      if (listA != listB) h = true;
      else if (listA) h = (d.flags[p] & F_INSERT) != 0;
      else {
        h = d.keyStrLen[p] != d.keyStrLen[p - 1];
        for (u32 i = 0; !h && i < d.keyStrLen[p]; i++) h = arena[d.keyStrOff[p] + i] != arena[d.keyStrOff[p - 1] + i];
      }
    }
    head[p] = h ? 1u : 0u;
  }
};
struct GroupStatsKernel {   // group id = inclusive scan of heads - 1; counts rows and visible rows per group
  const u32* headScan /* exclusive scan of head */; const u32* head; const u32* succCnt; DocRows d; u32* groupOf; u32* groupRows; u32* groupVisible; u32* groupFirst; u64* errWord; int allowCounters;
  HD void operator()(size_t p) const {
    const u32 g = headScan[p] + head[p] - 1; groupOf[p] = g;
    atomic_add(&groupRows[g], 1u);
    if (succCnt[p] == 0) atomic_add(&groupVisible[g], 1u);
    if (head[p]) groupFirst[g] = (u32)p;
    if (flags_action(d.flags[p]) == ACT_INC || ((d.valLen[p] & 15) == 8 && flags_action(d.flags[p]) == ACT_SET && succCnt[p] > 0)) { if (!allowCounters) raise(errWord, KE_UNSUPPORTED_OP, p); }
  }
};

// ---------------------------------------------------------------- incremental: touched groups / objects
struct TouchKernel {   // new rows and the targets of new succ entries touch their group and object
  DocRows d; const u32* groupOf; const u32* firstNewSucc; u32* groupTouched; u32* objTouchedAt /* per position of the object's make row */;
  const u32* objPos /* per position: position of the object's make row or ROW_NONE (root) */; u32* rootTouched;
  HD void operator()(size_t p) const {
    if (d.time[p] == 0 && firstNewSucc[p] == 0xffffffffu) return;
    groupTouched[groupOf[p]] = 1;
    if (objPos[p] == ROW_NONE) *rootTouched = 1; else objTouchedAt[objPos[p]] = 1;
  }
};
// setupPatches: a touched object links itself into its parent (the group of its make row), recursively
struct LinkKernel {
  DocRows d; const u32* groupOf; const u32* groupVisible; const u32* objPos; u32* groupLinked; u32* objTouchedAt; u32* rootTouched; u32* linkDone; u32* changed; u64* errWord;
  HD void operator()(size_t p) const {
    if (!objTouchedAt[p] || linkDone[p]) return;
    linkDone[p] = 1; *changed = 1;
    const u32 g = groupOf[p];
    if (groupVisible[g] == 0) return;   // hasChildren false: nothing to link (new.js:1465,1521)
    if (d.keyStrLen[p] == NULL32) { raise(errWord, KE_UNSUPPORTED_OP, p); return; }   // child object inside a list
    groupLinked[g] = 1;
    if (objPos[p] == ROW_NONE) *rootTouched = 1; else objTouchedAt[objPos[p]] = 1;
  }
};
// ---------------------------------------------------------------- which conflicting values the reference re-emits
// mergeDocChangeOps (new.js:1085-1138) processes the change ops of one author in "groups" (consecutive ops on the
// same key that do not overwrite each other). After a group's last op is placed, the remaining document ops of
// that key (those with a greater opId) are only re-emitted into the patch if no further group could be gathered
// into the same pass (new.js:1119-1129, 1149); otherwise they are passed over silently (new.js:1225-1230). The
// patch content of a key is what the LAST pass touching it emitted (props[key] is reset per pass, new.js:1037).
HD int key_cmp_utf16(const u8* a, u32 la, const u8* b, u32 lb) {   // JS string `<` on UTF-8 bytes
  const u32 n = la < lb ? la : lb;
  for (u32 i = 0; i < n; i++) {
    u32 x = a[i], y = b[i];
    if (x != y) {
      if (x == 0xEE || x == 0xEF) x += 5; else if (x >= 0xF0 && x <= 0xF4) x -= 2;
      if (y == 0xEE || y == 0xEF) y += 5; else if (y >= 0xF0 && y <= 0xF4) y -= 2;
      return x < y ? -1 : 1;
    }
  }
  return la < lb ? -1 : (la > lb ? 1 : 0);
}
struct OpAtTimeKernel { const u32* time; u32* opAt; HD void operator()(size_t i) const { opAt[time[i] - 1] = (u32)i; } };
struct MapGroupCtx {
  const u8* arena; OpRows ops; const u32* opAt; size_t numOps;
  HD bool isMapOp(u32 i) const { return ops.keyStrLen[i] != NULL32; }
  HD bool sameKey(u32 i, u32 j) const {
    if (ops.keyStrLen[i] != ops.keyStrLen[j]) return false;
    for (u32 k = 0; k < ops.keyStrLen[i]; k++) if (arena[ops.keyStrOff[i] + k] != arena[ops.keyStrOff[j] + k]) return false;
    return true;
  }
  HD bool sameRun(u32 i, u32 j) const {
    return isMapOp(i) && isMapOp(j) && id_actor(ops.id[i]) == id_actor(ops.id[j]) && ((ops.flags[i] ^ ops.flags[j]) & F_INSERT) == 0 &&
           ops.obj[i] == ops.obj[j] && sameKey(i, j);
  }
};
struct RunHeadKernel { MapGroupCtx c; u32* runHead; HD void operator()(size_t t) const { runHead[t] = (t == 0 || !c.sameRun(c.opAt[t], c.opAt[t - 1])) ? 1u : 0u; } };
struct GroupSplitKernel {   // one thread per run: a new group starts where an op overwrites an op of the current group (new.js:1092-1101)
  MapGroupCtx c; const u32* runHead; u32* groupHead;
  HD void operator()(size_t t0) const {
    if (!runHead[t0]) return;
    size_t gstart = t0; groupHead[t0] = 1;
    for (size_t t = t0 + 1; t < c.numOps && !runHead[t]; t++) {
      const u32 i = c.opAt[t]; bool over = false;
      for (u32 j = 0; j < c.ops.predNum[i] && !over; j++) {
        const u64 p = c.ops.predId[c.ops.predOff[i] + j];
        for (size_t u = gstart; u < t && !over; u++) over = c.ops.id[c.opAt[u]] == p;
      }
      groupHead[t] = over ? 1u : 0u;
      if (over) gstart = t;
    }
  }
};
struct GroupFinalKernel {   // pass 0: finalTime[g] = latest group on key group g; pass 1: that group publishes bound / failed / members
  int pass; MapGroupCtx c; const u32* groupHead; IdTable t; const u32* rowOfOp; const u32* pos; const u32* groupOf; DocRows w; Ord ord;
  u32* finalTime; u64* bound; u32* failed; u32* member;
  HD void operator()(size_t t0) const {
    if (!groupHead[t0]) return;
    const u32 i0 = c.opAt[t0]; if (!c.isMapOp(i0)) return;
    u64 b = 0; u32 g = ROW_NONE; size_t t = t0;
    for (; t < c.numOps && (t == t0 || !groupHead[t]); t++) {
      const u32 i = c.opAt[t];
      if (flags_action(c.ops.flags[i]) == ACT_DEL) {
        for (u32 j = 0; j < c.ops.predNum[i]; j++) {
          const u32 target = id_lookup(this->t, c.ops.predId[c.ops.predOff[i] + j]); if (target == ROW_NONE) continue;
          const u64 o = ord(w.id[target]); if (o > b) b = o; g = groupOf[pos[target]];
        }
      } else {
        const u64 o = ord(c.ops.id[i]); if (o > b) b = o;
        if (rowOfOp[i] != ROW_NONE) { g = groupOf[pos[rowOfOp[i]]]; if (pass == 1 && finalTime[g] == (u32)t0 + 1) member[pos[rowOfOp[i]]] = 1; }
      }
    }
    if (g == ROW_NONE) return;
    if (pass == 0) { atomic_max(&finalTime[g], (u32)t0 + 1); return; }
    if (finalTime[g] != (u32)t0 + 1) return;
    bool gathered = false;
    if (t < c.numOps) {
      const u32 nx = c.opAt[t];
      gathered = c.isMapOp(nx) && id_actor(c.ops.id[nx]) == id_actor(c.ops.id[i0]) && ((c.ops.flags[nx] ^ c.ops.flags[i0]) & F_INSERT) == 0 && c.ops.obj[nx] == c.ops.obj[i0] &&
                 key_cmp_utf16(c.arena + c.ops.keyStrOff[i0], c.ops.keyStrLen[i0], c.arena + c.ops.keyStrOff[nx], c.ops.keyStrLen[nx]) < 0;
    }
    bound[g] = b; failed[g] = gathered ? 0u : 1u;
    // members of a group that straddles the walk above were flagged on the fly; flag again now that finalTime is known
    for (size_t u = t0; u < t; u++) { const u32 i = c.opAt[u]; if (rowOfOp[i] != ROW_NONE) member[pos[rowOfOp[i]]] = 1; }
  }
};
struct PropFlagKernel {   // which positions emit a prop record
  DocRows d; const u32* groupOf; const u32* groupTouched; const u32* groupLinked; const u32* succCnt; int wholeDoc;
  const u32* finalTime; const u64* bound; const u32* failed; const u32* member; Ord ord; u32* emit; u32* groupEmitted;
  HD void operator()(size_t p) const {
    u32 e = 0;
    if (d.keyStrLen[p] != NULL32 && succCnt[p] == 0) {
      const u32 g = groupOf[p];
      if (wholeDoc) e = 1;
      else if (groupTouched[g]) e = (finalTime[g] == 0 || failed[g] || member[p] || ord(d.id[p]) <= bound[g]) ? 1 : 0;
      else if (groupLinked[g]) e = 1;
    }
    emit[p] = e;
    if (e) atomic_add(&groupEmitted[groupOf[p]], 1u);
  }
};
struct PropMarkerKernel {   // a touched key with nothing to show is reported as `key: {}` (new.js:1037)
  DocRows d; const u32* groupOf; const u32* groupTouched; const u32* head; const u32* groupEmitted; int wholeDoc; u32* emit; u32* marker;
  HD void operator()(size_t p) const {
    u32 m = 0;
    if (!wholeDoc && d.keyStrLen[p] != NULL32 && head[p] && groupTouched[groupOf[p]] && groupEmitted[groupOf[p]] == 0) { m = 1; emit[p] = 1; }
    marker[p] = m;
  }
};
struct PropEmitKernel {
  DocRows d; const u32* emit; const u32* marker; const u32* slot; PropRec* out;
  HD void operator()(size_t p) const {
    if (!emit[p]) return;
    PropRec r; r.obj = d.obj[p]; r.opId = d.id[p]; r.keyOff = d.keyStrOff[p]; r.keyLen = d.keyStrLen[p]; r.valLen = d.valLen[p]; r.valOff = d.valOff[p];
    r.flags = (flags_action(d.flags[p]) << 8) | (marker[p] ? 1u : 0u); r.pad = 0; out[slot[p]] = r;
  }
};

// ---------------------------------------------------------------- whole-document list edits (getPatch)
struct ListVisFlagKernel {   // element visible (any visible row) flagged on the group head; visible rows flagged individually
  DocRows d; const u32* groupOf; const u32* groupVisible; const u32* head; const u32* succCnt; u32* elemVis; u32* rowEmit;
  HD void operator()(size_t p) const {
    const bool list = d.keyStrLen[p] == NULL32;
    elemVis[p] = (list && head[p] && groupVisible[groupOf[p]] > 0) ? 1u : 0u;
    rowEmit[p] = (list && succCnt[p] == 0) ? 1u : 0u;
  }
};
struct ObjHeadKernel { DocRows d; u32* isObjHead; HD void operator()(size_t p) const { isObjHead[p] = (p == 0 || d.obj[p] != d.obj[p - 1]) ? 1u : 0u; } };
struct ObjStartKernel {   // objIdx = exclusive scan of heads (+head-1 fix-up); objStart[k] = first position of the k-th object
  const u32* isObjHead; u32* objIdx /* in: exclusive scan of isObjHead, out: object index */; u32* objStart; size_t n;
  HD void operator()(size_t p) const {
    const u32 k = objIdx[p] + isObjHead[p] - 1; objIdx[p] = k;
    if (isObjHead[p]) objStart[k] = (u32)p;
  }
};
struct DocEditEmitKernel {   // getPatch: visible rows of list objects in document order
  DocRows d; const u32* rowEmit; const u32* slot; const u32* elemVisScan /* exclusive */; const u32* objIdx; const u32* objStart;
  const u32* groupOf; const u32* groupFirst; const u32* succCnt; const u32* firstVisInGroup /* per group: first visible position */; EditRec* out;
  HD void operator()(size_t p) const {
    if (!rowEmit[p]) return;
    const u32 g = groupOf[p]; const u32 gp = groupFirst[g];
    EditRec e; e.obj = d.obj[p]; e.opId = d.id[p];
    e.index = elemVisScan[gp] - elemVisScan[objStart[objIdx[p]]];
    const u32 kind = (firstVisInGroup[g] == (u32)p) ? EK_INSERT : EK_UPDATE;
    e.kind = kind | (flags_action(d.flags[p]) << 16); e.valLen = d.valLen[p]; e.valOff = d.valOff[p];
    out[slot[p]] = e;
  }
};
struct FirstVisKernel { const u32* groupOf; const u32* succCnt; u32* firstVisInGroup; HD void operator()(size_t p) const { if (succCnt[p] == 0) atomic_min(&firstVisInGroup[groupOf[p]], (u32)p); } };
// elemId of an edit = id of the group's insert row (needed for insert edits whose opId differs, i.e. conflicts)
struct EditElemKernel { DocRows d; const u32* rowEmit; const u32* slot; const u32* groupOf; const u32* groupFirst; u64* elemOut; HD void operator()(size_t p) const { if (rowEmit[p]) elemOut[slot[p]] = d.id[groupFirst[groupOf[p]]]; } };

// ---------------------------------------------------------------- incremental list edits: dominance counting
struct DomItem { u32 time; u32 ref /* bit31: query; bits30..0: op index (query) */; int w; u32 acc; u32 gs, ge; };

// Builds items in (position) order. Per position p holding a list element's insert row:
//   queries first (the insert op itself if new; the first deleter if any), then points (+1 insert, -1 first deletion).
enum { ES_ELEM = 1, ES_VISIBLE_BEFORE = 2, ES_NEW = 4, ES_DEL_NOW = 8 };
struct ElemStateKernel {   // per position: state of the list element whose insert row sits there
  DocRows d; const u32* succCnt; const u32* newSuccCnt; const u32* firstNewSucc; const u32* groupRows; const u32* groupVisible; const u32* groupOf; const u32* groupTouched;
  u32* state; u32* nItems; u64* errWord;
  HD void operator()(size_t p) const {
    u32 st = 0, n = 0;
    const bool list = d.keyStrLen[p] == NULL32;
    if (list && (d.flags[p] & F_INSERT)) {
      const u32 g = groupOf[p]; st = ES_ELEM;
      const bool isNew = d.time[p] != 0, delNow = firstNewSucc[p] != 0xffffffffu;
      if (groupTouched[g] && (groupRows[g] != 1 || flags_action(d.flags[p]) != ACT_SET)) raise(errWord, KE_UNSUPPORTED_OP, p);
      const bool visibleBefore = !isNew && (groupRows[g] == 1 ? (succCnt[p] - newSuccCnt[p]) == 0 : groupVisible[g] > 0);
      if (isNew) st |= ES_NEW;
      if (visibleBefore) st |= ES_VISIBLE_BEFORE;
      if (delNow) st |= ES_DEL_NOW;
      const bool live = isNew || visibleBefore;      // contributes points during this call
      if (live) { n += 1; if (delNow) n += 1; }      // +1 insertion, -1 first deletion
      if (isNew) n += 1;                             // insert query
      if (delNow && live) n += 1;                    // remove query (first deleter wins, new.js:1026)
    } else if (list && (d.time[p] != 0 || firstNewSucc[p] != 0xffffffffu)) {
      raise(errWord, KE_UNSUPPORTED_OP, p);          // update of an existing list element in incremental mode
    }
    state[p] = st; nItems[p] = n;
  }
};
struct DomBuildKernel {
  DocRows d; const u32* state; const u32* firstNewSucc; const u32* itemBase; const u32* objIdx; const u32* objStart /* [numObjs+1] positions */;
  DomItem* items;
  HD void operator()(size_t p) const {
    const u32 st = state[p]; if (!(st & ES_ELEM)) return;
    const bool isNew = st & ES_NEW, delNow = st & ES_DEL_NOW, live = (st & ES_NEW) || (st & ES_VISIBLE_BEFORE);
    u32 k = itemBase[p];
    const u32 gs = itemBase[objStart[objIdx[p]]], ge = itemBase[objStart[objIdx[p] + 1]];
    if (isNew) { DomItem q; q.time = d.time[p]; q.ref = 0x80000000u | (u32)(2 * p); q.w = 0; q.acc = 0; q.gs = gs; q.ge = ge; items[k++] = q; }
    if (delNow && live) { DomItem q; q.time = firstNewSucc[p]; q.ref = 0x80000000u | (u32)(2 * p + 1); q.w = 0; q.acc = 0; q.gs = gs; q.ge = ge; items[k++] = q; }
    if (live) {
      DomItem a; a.time = d.time[p]; a.ref = 0; a.w = 1; a.acc = 0; a.gs = gs; a.ge = ge; items[k++] = a;
      if (delNow) { DomItem b; b.time = firstNewSucc[p]; b.ref = 0; b.w = -1; b.acc = 0; b.gs = gs; b.ge = ge; items[k++] = b; }
    }
  }
};
struct DomScanInput {   // per level: low word = 1 if the time bit is clear, high word = the item's weight if the bit is clear
  const DomItem* items; int bit;
  HD u64 operator()(size_t i) const {
    const bool z = ((items[i].time >> bit) & 1u) == 0;
    return z ? (1ull | ((u64)(u32)items[i].w << 32)) : 0ull;
  }
};
struct DomLevelKernel {   // accumulate + stable split of every group on `bit`; ZW = packed exclusive scan of DomScanInput
  const DomItem* in; DomItem* out; const u64* ZW; int bit;
  HD void operator()(size_t i) const {
    DomItem it = in[i];
    const u64 s_i = ZW[i], s_gs = ZW[it.gs], s_ge = ZW[it.ge];
    const u32 zg = (u32)s_ge - (u32)s_gs;             // zeros in the group
    const u32 zb = (u32)s_i - (u32)s_gs;               // zeros before i in the group
    const bool one = (it.time >> bit) & 1u;
    u32 dst;
    if (one) {
      if (it.ref & 0x80000000u) it.acc += (u32)(s_i >> 32) - (u32)(s_gs >> 32);
      dst = it.gs + zg + ((u32)i - it.gs - zb);
      it.gs = it.gs + zg;
    } else {
      dst = it.gs + zb;
      it.ge = it.gs + zg;
    }
    out[dst] = it;
  }
};
struct DomResultKernel {   // route query results back: qIndex[2p + which] = index
  const DomItem* items; u32* qIndex;
  HD void operator()(size_t i) const { if (items[i].ref & 0x80000000u) qIndex[items[i].ref & 0x7fffffffu] = items[i].acc; }
};
// edits in application order: one slot per op of the batch (insert rows -> insert edit, first deleters -> remove edit)
struct OpEditFlagKernel {
  OpRows ops; const u32* pos; IdTable t; const u32* firstNewSuccAtPos; const u32* stateAtPos; u32* emit;
  HD void operator()(size_t i) const {
    u32 e = 0;
    if (ops.keyStrLen[i] == NULL32) {
      const u32 act = flags_action(ops.flags[i]);
      if (act == ACT_DEL) {
        // emits a remove iff this op is the first deleter of a previously visible element
        for (u32 j = 0; j < ops.predNum[i] && !e; j++) {
          const u32 target = id_lookup(t, ops.predId[ops.predOff[i] + j]); if (target == ROW_NONE) continue;
          const u32 p = pos[target];
          if (firstNewSuccAtPos[p] == ops.time[i] && (stateAtPos[p] & (ES_NEW | ES_VISIBLE_BEFORE))) e = 1;
        }
      } else if (ops.flags[i] & F_INSERT) e = 1;
    }
    emit[i] = e;
  }
};
struct OpEditEmitKernel {
  OpRows ops; const u32* emit; const u32* slot; const u32* rowOfOp; const u32* pos; IdTable t; const u32* qIndex; EditRec* out; u64* elemOut; u32* objKeyOut /* sort key: index of the object in document order */; u32* timeOut; const u32* objIdx;
  HD void operator()(size_t i) const {
    if (!emit[i]) return;
    EditRec e; e.obj = ops.obj[i]; u32 p;
    if (flags_action(ops.flags[i]) == ACT_DEL) {
      u32 target = ROW_NONE;
      for (u32 j = 0; j < ops.predNum[i] && target == ROW_NONE; j++) target = id_lookup(t, ops.predId[ops.predOff[i] + j]);
      p = pos[target];
      e.opId = ops.id[i]; e.index = qIndex[2 * p + 1]; e.kind = EK_REMOVE | (ACT_DEL << 16); e.valLen = 0; e.valOff = 0; elemOut[slot[i]] = 0;
    } else {
      p = pos[rowOfOp[i]];
      e.opId = ops.id[i]; e.index = qIndex[2 * p]; e.kind = EK_INSERT | (flags_action(ops.flags[i]) << 16); e.valLen = ops.valLen[i]; e.valOff = ops.valOff[i]; elemOut[slot[i]] = ops.id[i];
    }
    out[slot[i]] = e; objKeyOut[slot[i]] = objIdx[p]; timeOut[slot[i]] = ops.time[i];
  }
};
// appendEdit coalescing (new.js:747-782): edit j continues the run of edit j-1
struct RunFlagKernel {
  EditRec* edits; const u64* elem; size_t n;
  HD u32 cls(u32 valLen) const { const u32 t = valLen & 15; return t == 2 ? 1 : t; }
  HD void operator()(size_t j) const {
    bool cont = false;
    if (j > 0) {
      const EditRec a = edits[j - 1], b = edits[j]; const u32 ka = a.kind & 0xff, kb = b.kind & 0xff;
      if (a.obj == b.obj) {
        if (ka == EK_INSERT && kb == EK_INSERT) {
          const u32 actA = (a.kind >> 16) & 0xffff, actB = (b.kind >> 16) & 0xffff;
          cont = b.index == a.index + 1 && actA == ACT_SET && actB == ACT_SET && elem[j - 1] == a.opId && elem[j] == b.opId &&
                 id_actor(a.opId) == id_actor(b.opId) && id_ctr(a.opId) + 1 == id_ctr(b.opId) && cls(a.valLen) == cls(b.valLen);
        } else if (ka == EK_REMOVE && kb == EK_REMOVE) cont = a.index == b.index;
      }
    }
    if (!cont) edits[j].kind |= 0x100u;   // run start
  }
};

}  // namespace amg
