// amgpu — list index, last levels of the dominance count inside shared memory (see DomLevelKernel in patch.cuh).
#pragma once
#include "patch.cuh"
#include "prims.cuh"

namespace amg {
#ifndef AMG_EMU
struct DomLocalBuf { u32 tw[DOM_LOCAL_MAX]; int w[DOM_LOCAL_MAX]; u32 acc[DOM_LOCAL_MAX]; unsigned short gs[DOM_LOCAL_MAX], ge[DOM_LOCAL_MAX]; };
struct DomLocalSmem { DomLocalBuf buf[2]; unsigned short Z[DOM_LOCAL_MAX + 2]; int W[DOM_LOCAL_MAX + 2]; u64 scan[9]; };
__global__ void __launch_bounds__(256) k_dom_local(const DomItem* __restrict__ items, const u32* __restrict__ partHead, const u32* __restrict__ numPartsPtr, u32* __restrict__ qIndex, int bits, u64* errWord) {
  extern __shared__ __align__(16) unsigned char domSmemRaw[];
  DomLocalSmem& S = *reinterpret_cast<DomLocalSmem*>(domSmemRaw);
  const u32 numParts = *numPartsPtr; const int IT = DOM_LOCAL_MAX / 256;
  for (u32 part = blockIdx.x; part < numParts; part += gridDim.x) {
    const u32 g0 = partHead[part]; const u32 m = items[g0].ge - g0;
    if (m > (u32)DOM_LOCAL_MAX) { if (threadIdx.x == 0) raise(errWord, KE_TOO_LARGE, g0); continue; }   // cannot happen (<= 2 items per time)
    for (u32 i = threadIdx.x; i < m; i += 256) { const DomItem it = items[g0 + i]; S.buf[0].tw[i] = it.tw; S.buf[0].w[i] = it.w; S.buf[0].acc[i] = it.acc; S.buf[0].gs[i] = 0; S.buf[0].ge[i] = (unsigned short)m; }
    __syncthreads();
    int cur = 0;
    for (int bit = bits - 1; bit >= 0; bit--) {
      DomLocalBuf& A = S.buf[cur]; DomLocalBuf& B = S.buf[cur ^ 1];
      // exclusive prefix over the items of (bit clear ? 1 : 0) and (bit clear ? weight : 0)
      const u32 lo = threadIdx.x * IT; u64 sum = 0;
      for (int k = 0; k < IT; k++) { const u32 i = lo + k; if (i < m && ((A.tw[i] >> bit) & 1u) == 0) sum += 1ull | ((u64)(u32)A.w[i] << 32); }
      u64 total; u64 ex = block_excl_scan64(sum, &total, S.scan);
      if (threadIdx.x == 0) { S.Z[m] = (unsigned short)(u32)total; S.W[m] = (int)(u32)(total >> 32); }
      for (int k = 0; k < IT; k++) {
        const u32 i = lo + k; if (i >= m) break;
        S.Z[i] = (unsigned short)(u32)ex; S.W[i] = (int)(u32)(ex >> 32);
        if (((A.tw[i] >> bit) & 1u) == 0) ex += 1ull | ((u64)(u32)A.w[i] << 32);
      }
      __syncthreads();
      for (u32 i = threadIdx.x; i < m; i += 256) {
        const u32 tw = A.tw[i]; const u32 gs = A.gs[i], ge = A.ge[i];
        const u32 zg = (u32)S.Z[ge] - (u32)S.Z[gs], zb = (u32)S.Z[i] - (u32)S.Z[gs];
        u32 acc = A.acc[i], dst, ngs = gs, nge = ge;
        if ((tw >> bit) & 1u) { if (dom_query(tw)) acc += (u32)(S.W[i] - S.W[gs]); dst = gs + zg + (i - gs - zb); ngs = gs + zg; }
        else { dst = gs + zb; nge = gs + zg; }
        B.tw[dst] = tw; B.w[dst] = A.w[i]; B.acc[dst] = acc; B.gs[dst] = (unsigned short)ngs; B.ge[dst] = (unsigned short)nge;
      }
      __syncthreads();
      cur ^= 1;
    }
    for (u32 i = threadIdx.x; i < m; i += 256) if (dom_query(S.buf[cur].tw[i])) qIndex[dom_time(S.buf[cur].tw[i]) - 1] = S.buf[cur].acc[i];
    __syncthreads();
  }
}
#endif

}  // namespace amg
