// Per-cell physics of the fluctuating binary (two-component Shan-Chen type) D3Q19 update.
// What is computed follows the reference cell by cell (algebra in SURVEY.md 3.2b):
//   hydrodynamic fields  LBM_binary.H:196-295 (hydrovars), :315-340 (hydrovars_bar_density)
//   19-point gradient    LBM_binary.H:134-150
//   noise amplitudes     LBM_binary.H:73-132
//   equilibrium / force moments and the relaxation  LBM_binary.H:356-516
// How it is computed is GPU-first: three reciprocals per cell instead of ~80 divisions, moments
// m0..m3 reused for densities and momenta, equilibrium and forcing written directly in their
// simplified algebraic form.  Agreement with the reference is at rounding level (tests: 1e-12).
#pragma once
#include <cfloat>
#include "d3q19.cuh"
#include "philox.cuh"

namespace bflbm {

struct DevParams {
  double rate_f, rate_g;  // 1/(tau (1 + 0.5/tau))            LBM_binary.H:504-505
  double fric_f, fric_g;  // 0.5/(tau + 0.5)                   LBM_binary.H:266-272
  double force_pf;        // 1/(1 + 1/(2 tau_f)), both species LBM_binary.H:424
  double acc_coef;        // -cs2 * alpha0                     LBM_binary.H:254-255
  double amp_j;           // A kBT,      A = 2(l - l^2/2), l = 1/(tau_f + 1/2)   LBM_binary.H:79-82,117
  double amp_s;           // A kBT / cs2                                          LBM_binary.H:125-126
  double sqrt_amp_j, sqrt_amp_s;  // their square roots (the amplitudes are sqrt(amp) * sqrt|density|)
  PhiloxKeys keys;        // Philox round keys of the seed (LBM_binary.H:17)
};

struct CellHydro {
  double rho, phi, inv_tot;
  double ufb[3], ugb[3];  // LB ("bar") velocities j/rho
  double af[3], ag[3];    // accelerations
  double uf[3], ug[3];    // real species velocities (contain +xi/2)
  double nfv[3], ngv[3];  // momentum noise / density
  double xi[3];           // momentum-mode noise of species f (species g gets -xi)
};

// (1/cs2) sum_i w_i n(x+c_i) c_i ; n[i] = field at x + c_i (n[0] unused)
__device__ __forceinline__ void gradient19(const double (&n)[Q], double (&g)[3]) {
  g[0] = (1. / 6.) * (n[1] - n[2]) + (1. / 12.) * (((n[7] - n[8]) + (n[9] - n[10])) + ((n[15] - n[16]) + (n[17] - n[18])));
  g[1] = (1. / 6.) * (n[3] - n[4]) + (1. / 12.) * (((n[7] - n[8]) - (n[9] - n[10])) + ((n[11] - n[12]) + (n[13] - n[14])));
  g[2] = (1. / 6.) * (n[5] - n[6]) + (1. / 12.) * (((n[11] - n[12]) - (n[13] - n[14])) + ((n[15] - n[16]) - (n[17] - n[18])));
}

// 1/sqrt(x) for x > 0 to ~1 ulp: hardware seed (rsqrt.approx.f64 = MUFU.RSQ64H, ~2^-22) + two Newton steps.
// No special-case code (the callers guard x = 0 and denormal x, which the .ftz seed flushes): 9 instructions instead of
// the ~15 + slow-path branch of sqrt() or 1/x.
__device__ __forceinline__ double rsqrt_pos(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  return y;
}

// y3: the first three standard normals of the cell (draws 0..2) in the generator's biased form (philox.cuh), ignored
// when !NOISE.  sq_rho / sq_phi return sqrt|rho|, sqrt|phi| (for the stress-mode noise amplitudes of collide_species).
// The three reciprocals (1/rho, 1/phi, 1/(rho+phi)) and three square roots the formulas need all come from three
// reciprocal square roots: 1/x = sign(x) r^2, sqrt|x| = |x| r with r = rsqrt|x|  (each ~2 ulp; the bar is 1e-12).
// ref3 != nullptr: USE_REF_STATE (LBM_binary.H:12, 92-107) -- the noise amplitudes are built from {rho, phi, rho_t} of the
// equilibrium profile at the COM-shifted cell instead of the current densities (thread-per-cell kernels only).
template <bool NOISE, bool FMA_NOISE = true>
__device__ __forceinline__ void cell_hydro(const DevParams& P, double rho, double phi, const double (&jf)[3], const double (&jg)[3],
                                           const double (&grad_rho)[3], const double (&grad_phi)[3], const float (&y3)[3],
                                           CellHydro& H, double& sq_rho, double& sq_phi, const double* ref3 = nullptr) {
  H.rho = rho;
  H.phi = phi;
  const bool has_f = fabs(rho) > (double)FLT_EPSILON, has_g = fabs(phi) > (double)FLT_EPSILON;
  const double tot = rho + phi;
  const double ar = fabs(rho), ap = fabs(phi), at = fabs(tot);
  // rsqrt.approx flushes denormals to zero and hands back inf, which the Newton steps turn into NaN.  A denormal total
  // density gives 1/(rho+phi) = inf here (the reference's division: > 4e307 or inf -- the run is lost either way); the
  // square root of a denormal density (< 1.5e-154) is taken as 0.
  const double rr = rsqrt_pos(ar), rp = rsqrt_pos(ap);
  const double rt = at >= DBL_MIN ? rsqrt_pos(at) : __longlong_as_double(0x7ff0000000000000ll);
  const double inv_rho = has_f ? copysign(rr * rr, rho) : 0., inv_phi = has_g ? copysign(rp * rp, phi) : 0.;
  H.inv_tot = copysign(rt * rt, tot);  // unguarded, like the reference (LBM_binary.H:266-272, 286-288, 471): tot = 0 -> inf
  sq_rho = ar >= DBL_MIN ? ar * rr : 0.;
  sq_phi = ap >= DBL_MIN ? ap * rp : 0.;
  // xi = sqrt(A kBT |rho phi / (rho + phi)|) n  (LBM_binary.H:117), amplitude = sqrt(A kBT) sqrt|rho| sqrt|phi| / sqrt|rho + phi|,
  // n = 16 (D - 1.5):  xi = fma(16 amp, D, -24 amp)
  double amp16 = 0., ampb = 0.;
  if (NOISE) {
    double amp = P.sqrt_amp_j * (sq_rho * sq_phi) * rt;
    if (ref3 != nullptr) {
      sq_rho = sqrt(fabs(ref3[0]));
      sq_phi = sqrt(fabs(ref3[1]));
      amp = P.sqrt_amp_j * sqrt(fabs(ref3[0] * ref3[1] / ref3[2]));
    }
    amp16 = NRM_SCALE * amp;
    ampb = -(NRM_SCALE * NRM_BIAS) * amp;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    H.ufb[k] = jf[k] * inv_rho;
    H.ugb[k] = jg[k] * inv_phi;
    H.af[k] = has_f ? P.acc_coef * grad_phi[k] : 0.;
    H.ag[k] = has_g ? P.acc_coef * grad_rho[k] : 0.;
    H.xi[k] = NOISE ? (FMA_NOISE ? fma(amp16, widen_pos(y3[k]), ampb) : (amp16 * (1. / NRM_SCALE)) * widen(normal_f32(y3[k]))) : 0.;
    H.nfv[k] = H.xi[k] * inv_rho;
    H.ngv[k] = -H.xi[k] * inv_phi;
    const double d = (H.ufb[k] - H.ugb[k]) + 0.5 * (H.af[k] - H.ag[k]);
    H.uf[k] = H.ufb[k] + 0.5 * H.af[k] - P.fric_f * phi * H.inv_tot * d + 0.5 * H.nfv[k];
    H.ug[k] = H.ugb[k] + 0.5 * H.ag[k] + P.fric_g * rho * H.inv_tot * d + 0.5 * H.ngv[k];
  }
}

// Collision of one cell, split in two so that the caller can finish species f (inverse transform, store) before
// touching species g:
//   collide_prepare : hydro fields + barycentric velocity (needs the conserved moments of both species)
//   collide_species : the part of the post-collision state that does not depend on the incoming non-conserved moments
//
// The reference relaxes EVERY moment with the same rate w = 1/(tau + 1/2) (LBM_binary.H:504-511):
//     m_a <- (1 - w) m_a + w meq_a + Phi_a + noise_a ,   a = 0..18
// and transforms back.  The transform is linear, so in population space
//     f_i <- (1 - w) f_i + [M^-1 (w meq + Phi + noise)]_i
// -- the forward transform is only needed for the conserved moments (densities and momenta) that meq, Phi and the noise
// amplitudes are built from, for ANY tau, and the old populations enter through one fma each.  collide_species returns
// v = w meq + Phi + noise; the caller adds (1 - w) f_old after the inverse transform.  With the reference's shipped
// tau = 1/2 (LBM_binary.H:18-19) w = 1 and that term vanishes (RATE1 kernels never keep f_old).
struct CollideCtx {
  CellHydro H;
  double vb[3];
  double sq_rho, sq_phi;  // sqrt|rho|, sqrt|phi|
};

// The standard normals come from the caller (momentum_normals / mode_normals below) in biased form: y3 = the three
// momentum-mode draws, yb = the 15 draws of modes 4..18 of the species.  Generating them is pure arithmetic on the cell's
// counter, so the step kernel does it in the shadow of its population loads, before the first loaded value is needed.
template <bool NOISE, bool FMA_NOISE = true>
__device__ __forceinline__ void collide_prepare(const DevParams& P, const double (&grad_rho)[3], const double (&grad_phi)[3],
                                                const float (&y3)[3], const double (&mf)[Q], const double (&mg)[Q], CollideCtx& C,
                                                const double* ref3 = nullptr) {
  const double jf[3] = {mf[1], mf[2], mf[3]}, jg[3] = {mg[1], mg[2], mg[3]};
  cell_hydro<NOISE, FMA_NOISE>(P, mf[0], mg[0], jf, jg, grad_rho, grad_phi, y3, C.H, C.sq_rho, C.sq_phi, ref3);
#pragma unroll
  for (int k = 0; k < 3; ++k) C.vb[k] = (C.H.rho * C.H.uf[k] + C.H.phi * C.H.ug[k]) * C.H.inv_tot;  // LBM_binary.H:471
}

// SPECIES 0 = f (call first), 1 = g.  v[0..18] = w meq(D, vb) + Phi(D, u, a) + noise (equilibrium LBM_binary.H:356-402,
// forcing :404-449, noise :113-127).  rate = w; RATE1: w is exactly 1 (saves the multiplications by it).
// INPLACE: v holds ALL 19 incoming moments and is relaxed in place, v_k <- (1 - w) v_k + (w meq + Phi + noise)_k -- the
// moment-space form of the same update (one array instead of two: used where registers are short).
// FMA_NOISE = false: the noise of modes 4..18 is added afterwards as  v_k += amplitude_k * widen(n_k)  (see philox.cuh).
template <bool NOISE, int SPECIES, bool RATE1, bool INPLACE = false, bool FMA_NOISE = true>
__device__ __forceinline__ void collide_species(const DevParams& P, const float (&yb)[15], const CollideCtx& C, double (&v)[Q]) {
  const CellHydro& H = C.H;
  const double D = SPECIES == 0 ? H.rho : H.phi, rate = SPECIES == 0 ? P.rate_f : P.rate_g;
  const double (&u)[3] = SPECIES == 0 ? H.uf : H.ug;
  const double (&a)[3] = SPECIES == 0 ? H.af : H.ag;
  const double (&vb)[3] = C.vb;
  const double Dr = RATE1 ? D : D * rate, Dp = D * P.force_pf;
  // noise of mode k: sqrt(A kBT/cs2 b_k |D|) n = fma(amp16 sb_k, Dn, -1.5 amp16 sb_k), Dn = widen_pos(y)
  const double s16 = NOISE ? NRM_SCALE * (P.sqrt_amp_s * (SPECIES == 0 ? C.sq_rho : C.sq_phi)) : 0.;
  const double keep = 1. - rate;
  auto old = [&](int k, double x) { return INPLACE ? fma(keep, v[k], x) : x; };
  auto set = [&](int k, double eq, double force) {  // v_k = Dr eq + Dp force + noise_k
    if (NOISE && FMA_NOISE) {
      const double amp = sqrt_bnorm(k) * s16;
      v[k] = fma(amp, widen_pos(yb[k - 4]), fma(Dr, eq, fma(Dp, force, old(k, -NRM_BIAS * amp))));
    } else {
      v[k] = fma(Dr, eq, old(k, Dp * force));
    }
  };
  v[0] = INPLACE ? D : Dr;  // keep D + w D
#pragma unroll
  for (int k = 0; k < 3; ++k) v[1 + k] = fma(Dr, vb[k], old(1 + k, Dp * a[k])) + (SPECIES == 0 ? H.xi[k] : -H.xi[k]);
  const double vxx = vb[0] * vb[0], vyy = vb[1] * vb[1], vzz = vb[2] * vb[2];
  const double axx = a[0] * u[0], ayy = a[1] * u[1], azz = a[2] * u[2];
  const double tr = axx + ayy + azz;
  set(4, vxx + vyy + vzz, 2. * tr);
  set(5, 2. * vxx - vyy - vzz, 6. * axx - 2. * tr);
  set(6, vyy - vzz, 2. * (ayy - azz));
  set(7, vb[0] * vb[1], a[0] * u[1] + a[1] * u[0]);
  set(8, vb[1] * vb[2], a[1] * u[2] + a[2] * u[1]);
  set(9, vb[0] * vb[2], a[0] * u[2] + a[2] * u[0]);
  // ghost modes: no equilibrium, no force
#pragma unroll
  for (int k = 10; k < Q; ++k) {
    if (NOISE && FMA_NOISE) {
      const double amp = sqrt_bnorm(k) * s16;
      v[k] = fma(amp, widen_pos(yb[k - 4]), old(k, -NRM_BIAS * amp));
    } else {
      v[k] = INPLACE ? keep * v[k] : 0.;
    }
  }
  if (NOISE && !FMA_NOISE) {
    const double s = s16 * (1. / NRM_SCALE);  // sqrt(A kBT/cs2 |D|), LBM_binary.H:125-126
#pragma unroll
    for (int k = 4; k < Q; ++k) v[k] += (sqrt_bnorm(k) * s) * widen(normal_f32(yb[k - 4]));
  }
}
// (1 - w) of the species: the weight of the old populations in the post-collision state
__device__ __forceinline__ double keep_of(const DevParams& P, int species) { return 1. - (species == 0 ? P.rate_f : P.rate_g); }

template <bool NOISE, bool TAB = false>
__device__ __forceinline__ void momentum_normals(const NoiseKey& nk, float (&y3)[3], const float2* __restrict__ trig = nullptr) {
  float y0[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (NOISE) species_normals<0, 0, 4, TAB>(nk, y0, trig);
  y3[0] = y0[0]; y3[1] = y0[1]; y3[2] = y0[2];
}
template <bool NOISE, int SPECIES, bool TAB = false>
__device__ __forceinline__ void mode_normals(const NoiseKey& nk, float (&yb)[15], const float2* __restrict__ trig = nullptr) {
  if (NOISE) {
    // modes 4..18 : normals F[3..17] (the pair F[2], F[3] comes from block 0, shared with the momentum draws) / G[0..14]
    constexpr int LO = SPECIES == 0 ? 2 : 0, HI = SPECIES == 0 ? 18 : 15;
    float t[HI - LO + 1];
    species_normals<SPECIES, LO, HI, TAB>(nk, t, trig);
#pragma unroll
    for (int a = 4; a < Q; ++a) yb[a - 4] = t[mode_index(SPECIES, a) - LO];
  } else {
#pragma unroll
    for (int j = 0; j < 15; ++j) yb[j] = 0.f;
  }
}

// the 33 standard normals of a cell in REFERENCE draw order (LBM_binary.H:115-127: a = 1..3 one draw each,
// a = 4..18 two draws each, f then g), biased form -- observer / test hook
__device__ __forceinline__ void cell_normals(const NoiseKey& nk, float (&y)[36]) {
  float F[19], G[16];
  species_normals<0, 0, 18>(nk, F);
  species_normals<1, 0, 15>(nk, G);
#pragma unroll
  for (int d = 0; d < 3; ++d) y[d] = F[d];
#pragma unroll
  for (int a = 4; a < Q; ++a) {
    y[3 + 2 * (a - 4)] = F[mode_index(0, a)];
    y[4 + 2 * (a - 4)] = G[mode_index(1, a)];
  }
  y[33] = y[34] = y[35] = NRM_BIAS_F;
}

}  // namespace bflbm
