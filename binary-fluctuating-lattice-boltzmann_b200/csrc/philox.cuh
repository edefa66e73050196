// Counter-based Gaussian noise for the fluctuating collision.
//
// The reference draws 33 N(0,1) numbers per cell per step from amrex::RandomNormal
// (LBM_binary.H:115-127): a=1..3 one draw each (shared by f and g with opposite sign), a=4..18 two
// draws each (f then g).  Here draw d in that same order is a pure function of
//     (seed, global cell index, step, d)
// through Philox4x32-10 (Salmon et al., SC'11) and a Box-Muller transform evaluated in fp32 with the
// hardware fast paths (MUFU lg2/sin/cos); the amplitude multiply is done in fp64 by the caller.
// Results therefore do not depend on the decomposition (number of GPUs, tiling) or launch order.
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

// two uniforms -> two independent standard normals
__device__ __forceinline__ void box_muller(uint32_t u0, uint32_t u1, float& n0, float& n1) {
  // U in (0,1]: (u0 + 0.5) / 2^32, tail down to 2^-33 (|n| <= 6.76)
  const float U = fmaf((float)u0, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  const float r = fast_sqrt(-1.3862943611198906f * fast_lg2(U));     // sqrt(-2 ln U), ln U = ln2 * lg2 U
  const float th = (float)(int32_t)u1 * 1.4629180792671596e-9f;      // pi * 2^-31 * s32  in [-pi, pi)
  n0 = r * fast_cos(th);
  n1 = r * fast_sin(th);
}

// Philox counter layout: {cell_lo, cell_hi, step_lo, (step_hi & 0xffffff) | block << 24}, key = seed.
// Block j yields the internal normals 4j .. 4j+3.  Internal order (species-symmetric, so that the two threads
// that share a cell in the fused kernel run identical code on different blocks):
//   0..2    momentum modes a = 1..3 (species f gets +xi, species g gets -xi); 3 unused
//   4..18   species f, modes a = 4..18 (blocks 1-4);  19 unused
//   20..34  species g, modes a = 4..18 (blocks 5-8);  35 unused
// draw_index() maps the reference's draw order (a = 4..18: f then g, interleaved) onto it.
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
// internal index of the normal that drives mode a (4..18) of species s (0 = f, 1 = g)
__host__ __device__ constexpr int mode_index(int s, int a) { return 4 + 16 * s + (a - 4); }
// reference draw d (0..32, LBM_binary.H:115-127) -> internal normal index
__host__ __device__ constexpr int draw_index(int d) { return d < 3 ? d : mode_index((d - 3) & 1, 4 + ((d - 3) >> 1)); }

__device__ __forceinline__ void normals4(const NoiseKey& k, int block, float (&n)[4]) {
  const uint4 r = philox4x32(make_uint4(k.cell_lo, k.cell_hi, k.step_lo, k.step_hi | ((uint32_t)block << 24)), *k.K);
  box_muller(r.x, r.y, n[0], n[1]);
  box_muller(r.z, r.w, n[2], n[3]);
}

}  // namespace bflbm
