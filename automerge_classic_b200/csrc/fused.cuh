// amgpu — kernels #1c: header parse and column expansion in ONE pass over the change bytes.
//
// The two-kernel decode (ParseKernel, then DecodeSmallKernel after the causal gate and a prefix sum of the op counts)
// reads every change twice and round-trips an 88-byte record per change through HBM. Here one kernel parses a change,
// obtains its row offset from a single-pass prefix sum (decoupled look-back over 256-change tiles, tickets handed out
// by an atomic counter so that a tile's predecessors are always running or done) and expands its columns right away.
// The offsets are those of "every change of the batch is applied" — true for the usual bulk replay; the causal gate
// runs afterwards and, if it drops or defers changes, the engine falls back to DecodeSmallKernel with gate-aware offsets
// (speculative rows are simply overwritten). Column errors found here are kept apart and only count in the first case.
#pragma once
#include "decode.cuh"

namespace amg {

static const u64 TS_AGG = 1ull << 62, TS_INCL = 2ull << 62, TS_MASK = 3ull << 62;
HD u64 ts_pack(u32 ops, u32 preds) { return ((u64)ops << 31) | (u64)preds; }   // both < 2^31 (the engine caps a document at 2^29 rows)
HD u32 ts_ops(u64 w) { return (u32)((w & ~TS_MASK) >> 31); }
HD u32 ts_preds(u64 w) { return (u32)(w & 0x7fffffffu); }

struct FusedArgs {
  ParseKernel parse; DecodeSmallKernel dec;   // dec.opBase / predBase / applied are not used here
  u32* opBaseOut; u32* predBaseOut;           // [n + 1]: exclusive prefix sums of nOps / nPreds in batch order
  u64* tileState; u32* ticket; u64* specErr; u32* overflow; u32 capOps, capPreds;
};

#ifndef AMG_EMU
__global__ void __launch_bounds__(256) k_parse_decode(size_t n, FusedArgs a) {
  __shared__ u32 sTile; __shared__ u64 sWarp[8]; __shared__ u64 sPrefix;
  const u8* arena = a.parse.arena;
  while (true) {
    if (threadIdx.x == 0) sTile = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const u32 tile = sTile; const size_t start = (size_t)tile * 256;
    if (start >= n) return;
    const size_t i = start + threadIdx.x;
    const u8* base = arena;   // (staging the tile's bytes in shared memory was measured: slower than reading through L1)
    u32 nOps = 0, nPreds = 0;
    if (i < n) { a.parse(i, base); nOps = a.parse.nOpsOut[i]; nPreds = a.parse.nPredsOut[i]; }
    // block exclusive scan of the packed counts
    u64 v = ts_pack(nOps, nPreds), incl = v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u64 t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) sWarp[warp] = incl;
    __syncthreads();
    u64 warpBase = 0, total = 0;
    for (int w = 0; w < 8; w++) { if (w < warp) warpBase += sWarp[w]; total += sWarp[w]; }
    if (threadIdx.x == 0) {   // publish the aggregate, look back for the exclusive prefix, publish the inclusive prefix
      volatile u64* ts = a.tileState; u64 excl = 0;
      if (tile == 0) ts[0] = TS_INCL | total;
      else {
        ts[tile] = TS_AGG | total;
        for (u32 p = tile; p-- > 0;) {
          u64 w; do { w = ts[p]; } while ((w & TS_MASK) == 0);
          excl += w & ~TS_MASK;
          if ((w & TS_MASK) == TS_INCL) break;
        }
        ts[tile] = TS_INCL | (excl + total);
      }
      sPrefix = excl;
    }
    __syncthreads();
    const u64 mine = sPrefix + warpBase + incl - v;   // exclusive prefix of this change
    if (i < n) {
      const u32 ob = ts_ops(mine), pb = ts_preds(mine);
      a.opBaseOut[i] = ob; a.predBaseOut[i] = pb;
      if (i == n - 1) { a.opBaseOut[n] = ob + nOps; a.predBaseOut[n] = pb + nPreds; }
      if (nOps > 0 && nOps <= SMALL_CHANGE_OPS) {
        if ((u64)ob + nOps <= a.capOps && (u64)pb + nPreds <= a.capPreds) a.dec.decodeAt(i, base, ob, pb, a.specErr);
        else *a.overflow = 1;
      } else if (nOps > SMALL_CHANGE_OPS && ((u64)ob + nOps > a.capOps || (u64)pb + nPreds > a.capPreds)) *a.overflow = 1;
    }
    __syncthreads();   // sTile / sWarp / sPrefix are reused by the next tile
  }
}
#endif

inline void parse_decode(Ctx& c, size_t n, const FusedArgs& a) {
  if (n == 0) return;
#ifdef AMG_EMU
  u32 ob = 0, pb = 0;
  for (size_t i = 0; i < n; i++) {
    a.parse(i, a.parse.arena);
    const u32 nOps = a.parse.nOpsOut[i], nPreds = a.parse.nPredsOut[i];
    a.opBaseOut[i] = ob; a.predBaseOut[i] = pb;
    if (nOps > 0 && nOps <= SMALL_CHANGE_OPS) {
      if ((u64)ob + nOps <= a.capOps && (u64)pb + nPreds <= a.capPreds) a.dec.decodeAt(i, a.parse.arena, ob, pb, a.specErr); else *a.overflow = 1;
    } else if (nOps > SMALL_CHANGE_OPS && ((u64)ob + nOps > a.capOps || (u64)pb + nPreds > a.capPreds)) *a.overflow = 1;
    ob += nOps; pb += nPreds;
  }
  a.opBaseOut[n] = ob; a.predBaseOut[n] = pb;
#else
  const size_t tiles = (n + 255) / 256, maxGrid = (size_t)c.numSMs * 8;
  const int grid = (int)(tiles < maxGrid ? tiles : maxGrid);
  k_parse_decode<<<grid, 256, 0, c.stream>>>(n, a);
  CUDA_CHECK(cudaGetLastError());
#endif
  c.launches++;
}

}  // namespace amg
