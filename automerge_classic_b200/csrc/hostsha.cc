// amgpu — SHA-256 on the host (FIPS 180-4) for the two container checksums that are one long serial hash: the document
// chunk in Backend.load / Backend.save (reference backend/columnar.js:674-679, 698-705; fast-sha256 there). Change
// hashes are computed on the device (decode.cuh ShaKernel). Uses the x86 SHA extensions when the CPU has them
// (about 5x the portable code), checked at run time.
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

const uint32_t K[64] = {
  0x428a2f98,0x71374491,0xb5c0fbcf,0xe9b5dba5,0x3956c25b,0x59f111f1,0x923f82a4,0xab1c5ed5,0xd807aa98,0x12835b01,0x243185be,0x550c7dc3,
  0x72be5d74,0x80deb1fe,0x9bdc06a7,0xc19bf174,0xe49b69c1,0xefbe4786,0x0fc19dc6,0x240ca1cc,0x2de92c6f,0x4a7484aa,0x5cb0a9dc,0x76f988da,
  0x983e5152,0xa831c66d,0xb00327c8,0xbf597fc7,0xc6e00bf3,0xd5a79147,0x06ca6351,0x14292967,0x27b70a85,0x2e1b2138,0x4d2c6dfc,0x53380d13,
  0x650a7354,0x766a0abb,0x81c2c92e,0x92722c85,0xa2bfe8a1,0xa81a664b,0xc24b8b70,0xc76c51a3,0xd192e819,0xd6990624,0xf40e3585,0x106aa070,
  0x19a4c116,0x1e376c08,0x2748774c,0x34b0bcb5,0x391c0cb3,0x4ed8aa4a,0x5b9cca4f,0x682e6ff3,0x748f82ee,0x78a5636f,0x84c87814,0x8cc70208,
  0x90befffa,0xa4506ceb,0xbef9a3f7,0xc67178f2};

inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

void blocks_portable(uint32_t h[8], const uint8_t* p, size_t nblocks) {
  for (; nblocks; nblocks--, p += 64) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
      const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
      const uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
      const uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
}

#if defined(__x86_64__)
// The SHA extensions keep the state as two registers {A,B,E,F} and {C,D,G,H}; sha256rnds2 does two rounds, sha256msg1 /
// sha256msg2 the two halves of the message schedule for four words at a time.
__attribute__((target("sha,sse4.1,ssse3")))
void blocks_shani(uint32_t h[8], const uint8_t* p, size_t nblocks) {
  const __m128i bswap = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
  __m128i t = _mm_loadu_si128((const __m128i*)&h[0]);    // D C B A (high .. low)
  __m128i s1 = _mm_loadu_si128((const __m128i*)&h[4]);   // H G F E
  t = _mm_shuffle_epi32(t, 0xB1);                        // C D A B
  s1 = _mm_shuffle_epi32(s1, 0x1B);                      // E F G H
  __m128i s0 = _mm_alignr_epi8(t, s1, 8);                // A B E F
  s1 = _mm_blend_epi16(s1, t, 0xF0);                     // C D G H
  for (; nblocks; nblocks--, p += 64) {
    const __m128i save0 = s0, save1 = s1;
    __m128i m[4];
    for (int i = 0; i < 16; i++) {
      if (i < 4) m[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(p + 16 * i)), bswap);
      else {
        __m128i x = _mm_sha256msg1_epu32(m[i & 3], m[(i + 1) & 3]);               // W[t-16] + s0(W[t-15])
        x = _mm_add_epi32(x, _mm_alignr_epi8(m[(i + 3) & 3], m[(i + 2) & 3], 4));  // + W[t-7]
        m[i & 3] = _mm_sha256msg2_epu32(x, m[(i + 3) & 3]);                        // + s1(W[t-2])
      }
      __m128i wk = _mm_add_epi32(m[i & 3], _mm_loadu_si128((const __m128i*)&K[4 * i]));
      s1 = _mm_sha256rnds2_epu32(s1, s0, wk);
      wk = _mm_shuffle_epi32(wk, 0x0E);
      s0 = _mm_sha256rnds2_epu32(s0, s1, wk);
    }
    s0 = _mm_add_epi32(s0, save0); s1 = _mm_add_epi32(s1, save1);
  }
  t = _mm_shuffle_epi32(s0, 0x1B);                       // F E B A
  s1 = _mm_shuffle_epi32(s1, 0xB1);                      // D C H G
  s0 = _mm_blend_epi16(t, s1, 0xF0);                     // D C B A
  s1 = _mm_alignr_epi8(s1, t, 8);                        // H G F E
  _mm_storeu_si128((__m128i*)&h[0], s0); _mm_storeu_si128((__m128i*)&h[4], s1);
}
bool have_shani() { static const bool v = !getenv("AMG_NO_SHANI") && __builtin_cpu_supports("sha") && __builtin_cpu_supports("sse4.1") && __builtin_cpu_supports("ssse3"); return v; }   // AMG_NO_SHANI: tests of the portable path
#endif

void blocks(uint32_t h[8], const uint8_t* p, size_t n) {
#if defined(__x86_64__)
  if (have_shani()) { blocks_shani(h, p, n); return; }
#endif
  blocks_portable(h, p, n);
}

}  // namespace

extern "C" void amg_host_sha256(const uint8_t* data, size_t len, uint8_t out[32]) {
  uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  const size_t full = len / 64;
  blocks(h, data, full);
  uint8_t tail[128]; const size_t rem = len - full * 64; memcpy(tail, data + full * 64, rem);
  tail[rem] = 0x80; const size_t padded = rem + 9 <= 64 ? 64 : 128; memset(tail + rem + 1, 0, padded - rem - 1);
  const uint64_t bits = (uint64_t)len * 8; for (int i = 0; i < 8; i++) tail[padded - 1 - i] = (uint8_t)(bits >> (8 * i));
  blocks(h, tail, padded / 64);
  for (int i = 0; i < 8; i++) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
}
extern "C" int amg_host_sha256_accelerated() {
#if defined(__x86_64__)
  return have_shani() ? 1 : 0;
#else
  return 0;
#endif
}
