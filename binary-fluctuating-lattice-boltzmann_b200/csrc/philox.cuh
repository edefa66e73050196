// Counter-based Gaussian noise for the fluctuating collision.
//
// The reference draws 33 N(0,1) numbers per cell per step from amrex::RandomNormal
// (LBM_binary.H:115-127): a=1..3 one draw each (shared by f and g with opposite sign), a=4..18 two
// draws each (f then g).  Here draw d in that same order is a pure function of
//     (seed, global cell index, step, d)
// through Philox4x32-10 (Salmon et al., SC'11) and a Box-Muller transform evaluated in fp32 with the
// hardware fast paths (MUFU lg2/sin/cos); the amplitude multiply is done in fp64 by the caller.
// Results therefore do not depend on the decomposition (number of GPUs, tiling) or launch order.
// One 32-bit Philox word makes one Box-Muller PAIR: 22 bits for the radius uniform (|n| <= 5.65, 4 M levels) and a 16-bit angle
// built from all 32 bits (angle_index below; the marginal of r cos(theta) over M equispaced angles differs from the continuous
// one only in Bessel terms J_M, J_2M, ...).  5 Philox blocks per cell (the first version spent 32 + 32 bits per pair: 9 blocks).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bflbm {

#ifndef BFLBM_PHILOX_ROUNDS
#define BFLBM_PHILOX_ROUNDS 10
#endif

__host__ __device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < BFLBM_PHILOX_ROUNDS; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * ctr.x, p1 = (uint64_t)0xCD9E8D57u * ctr.z;  // one IMAD.WIDE each
    ctr = make_uint4((uint32_t)(p1 >> 32) ^ ctr.y ^ key.x, (uint32_t)p1, (uint32_t)(p0 >> 32) ^ ctr.w ^ key.y, (uint32_t)p0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

// The same block function with the key schedule (key + r*W, identical for every thread) precomputed on the
// host and passed as kernel constants: the round keys become constant-bank operands of the xor.
struct PhiloxKeys {
  uint32_t k0[BFLBM_PHILOX_ROUNDS], k1[BFLBM_PHILOX_ROUNDS];
};
inline PhiloxKeys philox_key_schedule(unsigned long long seed) {
  PhiloxKeys K;
  uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
  for (int r = 0; r < BFLBM_PHILOX_ROUNDS; ++r) {
    K.k0[r] = a;
    K.k1[r] = b;
    a += 0x9E3779B9u;
    b += 0xBB67AE85u;
  }
  return K;
}
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, const PhiloxKeys& K) {
#pragma unroll
  for (int r = 0; r < BFLBM_PHILOX_ROUNDS; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * ctr.x, p1 = (uint64_t)0xCD9E8D57u * ctr.z;
    ctr = make_uint4((uint32_t)(p1 >> 32) ^ ctr.y ^ K.k0[r], (uint32_t)p1, (uint32_t)(p0 >> 32) ^ ctr.w ^ K.k1[r], (uint32_t)p0);
  }
  return ctr;
}

// MUFU fast paths, flush-to-zero, no denormal/IEEE fix-up code around them
__device__ __forceinline__ float fast_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_sin(float x) { float r; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_cos(float x) { float r; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// The generator hands out every standard normal n in BIASED form  y = 1.5 + n / 16  (float, always inside [1, 2)): the bias costs nothing (it is the addend of the last fma of the Box-Muller transform; |n| <= 5.65) and makes the
// float -> double widening two integer operations: sign 0 and exponent 0 are known, so the fp64 image of y is
//     hi = (bits >> 3) + 0x38000000   (exponent re-bias + top 20 mantissa bits),   lo = bits << 29
// -- exact.  The consumers never form n itself: amplitude * n = fma(16 amplitude, D, -24 amplitude) with D = widen_pos(y).
// History: cvt.f64.f32 goes through the low-rate conversion unit (SASS F2F) and cost 1.25 ms of a 17 ms step at 512^3; a
// general re-packing with sign handling is 5 ALU operations per normal, 165 per cell (round 1); this form is 66.
constexpr float NRM_BIAS_F = 1.5f;
constexpr double NRM_BIAS = 1.5;     // D = NRM_BIAS + n / NRM_SCALE
constexpr double NRM_SCALE = 16.;
__device__ __forceinline__ double widen_pos(float y) {  // exact for every positive normal float
  const unsigned u = __float_as_uint(y);
  return __hiloint2double((int)((u >> 3) + 0x38000000u), (int)(u << 29));
}
// General float -> double re-packing (sign handled: 5 integer operations), exact for every normal float, and the normal
// as an fp32 number: n = 16 y - 24 is exact in fp32 (y and 1.5 share their exponent).  The general-rate kernels, which are
// short of registers, apply the noise as  m += amplitude * widen(n)  (round 1's form, 18.1 ms against 19.2+ ms for the fma
// form at 512^3, profiles/README.md); both forms see the same normals.
__device__ __forceinline__ double widen(float f) {
  const unsigned u = __float_as_uint(f);
  const unsigned hi = (((u & 0x7fffffffu) >> 3) + 0x38000000u) | (u & 0x80000000u);
  return __hiloint2double((int)hi, (int)(u << 29));
}
__device__ __forceinline__ float normal_f32(float y) { return fmaf(y, 16.f, -24.f); }
// the standard normal itself (observers, tests): exact, y has 24 significant bits
__device__ __forceinline__ double normal_of(float y) { return NRM_SCALE * (widen_pos(y) - NRM_BIAS); }

// Angle of a word.  10 bits (bits 0..9, disjoint from the radius bits 10..31) put the pair (n0, n1) on one of 1024 rays: invisible
// in every marginal and in the pooled statistics of 1e10 normals, but a 2-D histogram of 2.8e9 pairs sees it (chi2/dof = 1.9 on
// 32 x 32 bins, tests/test_gpu_noise_quality.py).  16 bits: j = (64 w + (w >> 10)) mod 2^16 -- the coarse angle is bits 0..9
// plus radius bits (a one-time pad: uniform and independent of the radius), the fine angle is the radius's six lowest bits,
// which move the radius by < 2^-16 relative.  65 536 rays, every one of the 1024 coarse directions present at every radius level
// (the tails stay isotropic; taking the radius's low bits as the angle's HIGH bits does not: chi2 = 23 764 / 219 at 1e10,
// profiles/README.md).  One more integer instruction per pair than the 10-bit form.
#ifndef BFLBM_ANGLE_BITS
#define BFLBM_ANGLE_BITS 16
#endif
#ifndef BFLBM_TRIG_TABLE  // 1: the fused kernels read (cos, sin) from a 1024-entry shared-memory table (10-bit angle only)
#define BFLBM_TRIG_TABLE 0
#endif
#define BFLBM_TRIG_TABLE_ALLOWED BFLBM_TRIG_TABLE
constexpr uint32_t ANGLE_MASK = (1u << BFLBM_ANGLE_BITS) - 1u;
__device__ __forceinline__ uint32_t angle_index(uint32_t w) {
  return (BFLBM_ANGLE_BITS == 10 ? w : (w << 6) + (w >> 10)) & ANGLE_MASK;
}
constexpr float ANGLE_STEP = 6.283185307179586f / (float)(1u << BFLBM_ANGLE_BITS);  // 2 pi / 2^bits
// one Philox word -> two independent standard normals, biased form
__device__ __forceinline__ void box_muller(uint32_t w, float& y0, float& y1) {
  // U in (0,1): (w>>10 + 0.5) / 2^22 ; theta = 2 pi (angle bits + 0.5) / 2^bits ; both conversions are exact in fp32
  const float U = fmaf((float)(w >> 10), 2.384185791015625e-07f, 1.1920928955078125e-07f);
  // r / 16 = sqrt(-2 ln U) / 16 = sqrt(-(2 ln2 / 256) lg2 U).  lg2.approx of a U just below 1 may come out as +2^-22
  // instead of -0: the |.| (a free operand modifier of MUFU.SQRT) keeps the argument non-negative.
  const float r = fast_sqrt(fabsf(5.4152123481245727e-03f * fast_lg2(U)));
  const float th = fmaf((float)angle_index(w), ANGLE_STEP, 0.5f * ANGLE_STEP);
  y0 = fmaf(r, fast_cos(th), NRM_BIAS_F);
  y1 = fmaf(r, fast_sin(th), NRM_BIAS_F);
}

// The same pair with (cos, sin) of the word's angle read from a 1024-entry table in shared memory that was filled with exactly the
// two MUFU results above (fill_trig_table): a memoisation, bit-identical normals, 2 instructions (mask, LDS.64) instead of 7
// (mask, int->float, fma, 2 x (range-reduction multiply + MUFU)) per pair.
__device__ __forceinline__ void fill_trig_table(float2* tab, int tid, int nthreads) {
  static_assert(BFLBM_ANGLE_BITS == 10 || !BFLBM_TRIG_TABLE_ALLOWED, "the shared-memory trig table needs the 10-bit angle");
  for (int j = tid; j < 1024; j += nthreads) {
    const float th = fmaf((float)j, ANGLE_STEP, 0.5f * ANGLE_STEP);
    tab[j] = make_float2(fast_cos(th), fast_sin(th));
  }
}
__device__ __forceinline__ void box_muller(uint32_t w, const float2* __restrict__ tab, float& y0, float& y1) {
  const float U = fmaf((float)(w >> 10), 2.384185791015625e-07f, 1.1920928955078125e-07f);
  const float r = fast_sqrt(fabsf(5.4152123481245727e-03f * fast_lg2(U)));
  const float2 cs = tab[w & 1023u];
  y0 = fmaf(r, cs.x, NRM_BIAS_F);
  y1 = fmaf(r, cs.y, NRM_BIAS_F);
}

// Philox counter layout: {cell_lo, cell_hi, step_lo, (step_hi & 0xffffff) | block << 24}, key = seed.
// Word w (0..3) of block b yields the pair (cos, sin) = normals (2w, 2w+1) of that block.  Layout per cell:
//   species f : blocks 0, 1 and word 0 of block 2 -> 18 normals  F[0..2] momentum modes a = 1..3 (species g gets -xi),
//                                                                 F[3..17] modes a = 4..18
//   species g : blocks 3, 4                        -> 16 normals  G[0..14] modes a = 4..18
// so the two species never share a block (no block is computed twice, none has to stay live across a species).
struct NoiseKey {
  const PhiloxKeys* K;  // points at the kernel parameter (constant bank)
  uint32_t cell_lo, cell_hi, step_lo, step_hi;
};
__device__ __forceinline__ NoiseKey make_noise_key(const PhiloxKeys& K, unsigned long long cell, long long step) {
  NoiseKey k;
  k.K = &K;
  k.cell_lo = (uint32_t)cell;
  k.cell_hi = (uint32_t)(cell >> 32);
  k.step_lo = (uint32_t)(unsigned long long)step;
  k.step_hi = (uint32_t)((unsigned long long)step >> 32) & 0x00ffffffu;
  return k;
}
constexpr int NOISE_BLOCKS = 5;
// normal j (0..17) of species f / j (0..14) of species g -> (block, word, half)
__host__ __device__ constexpr int noise_block(int s, int j) { return (s ? 3 : 0) + (j >> 3); }
__host__ __device__ constexpr int noise_word(int j) { return (j >> 1) & 3; }
// index of the normal that drives mode a (4..18) of species s inside that species' list
__host__ __device__ constexpr int mode_index(int s, int a) { return s ? a - 4 : a - 1; }

__device__ __forceinline__ uint4 noise_block_words(const NoiseKey& k, int block) {
  return philox4x32(make_uint4(k.cell_lo, k.cell_hi, k.step_lo, k.step_hi | ((uint32_t)block << 24)), *k.K);
}
__device__ __forceinline__ uint32_t word_of(const uint4& r, int w) { return w == 0 ? r.x : (w == 1 ? r.y : (w == 2 ? r.z : r.w)); }

// normals lo .. hi-1 of species s into n[0 .. hi-lo-1] (lo even); compile-time bounds, everything unrolls
// TAB: trig is the shared-memory (cos, sin) table (compile-time switch: a pointer test would keep both code paths alive)
template <int S, int LO, int HI, bool TAB = false>
__device__ __forceinline__ void species_normals(const NoiseKey& k, float (&n)[HI - LO + 1], const float2* __restrict__ trig = nullptr) {
  static_assert((LO & 1) == 0, "pairs start at even indices");
  uint4 r = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int j = LO; j < HI; j += 2) {
    if (j == LO || (j & 7) == 0) r = noise_block_words(k, noise_block(S, j));
    float a, b;
    if (TAB) box_muller(word_of(r, noise_word(j)), trig, a, b);
    else     box_muller(word_of(r, noise_word(j)), a, b);
    n[j - LO] = a;
    n[j - LO + 1] = b;  // the array has one spare slot for an odd count
  }
}

}  // namespace bflbm
