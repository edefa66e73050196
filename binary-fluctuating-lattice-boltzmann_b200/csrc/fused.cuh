// One-pass fused step: pull-stream + hydro + collide (+noise) + store, AND the (rho, phi) field of the
// NEXT step, in a single sweep over the populations (608 B/cell of population traffic).
//
// Why: the collision at x needs grad rho, grad phi at x (LBM_binary.H:237-240, 254-255), i.e. post-stream
// densities at the 18 neighbours, which are themselves sums of populations streamed from THEIR neighbours:
// a radius-2 dependency hidden in the reference's step (hard part 1 of SURVEY.md section 7).  A second
// sweep over the populations would cost another 304 B/cell.  Instead each CTA scatter-accumulates the
// post-collision populations it has just produced into next-step densities:
//     rho_{n+1}(y) = sum_i f*_i(y - c_i)
//   - along x with warp shuffles, along y and z through a rolling 3-plane shared-memory accumulator;
//   - contributions that leave the CTA's brick go to the brick's private one-cell shell ("extended box") in E;
//   - k_fold then adds, in a fixed order, the <= 8 bricks whose extended boxes contain a cell.
// No atomics, no zero-fill pass, bit-reproducible, and independent of how many GPUs the box is cut into
// (as long as slab boundaries are brick boundaries).
//
// A CTA owns a brick of tx*ty columns and sweeps lz planes upward; the (rho, phi) neighbourhood needed
// for the gradients is staged through a second rolling 3-plane shared-memory tile.
#pragma once
#include "kernels.cuh"

namespace bflbm {

struct BrickGrid {
  int tx, ty, lz;   // brick = tx*ty threads (256), lz planes
  int bx, by, bz;   // bricks per axis (last one may be partial)
  int ex, ey;       // tx+2, ty+2
  int pl;           // ex*ey     : extended plane, in double2
  long long brick;  // pl*(lz+2) : extended box, in double2
};

inline BrickGrid make_brick_grid(const Geom& G, int lz_request = 0, int nthreads = 256, int max_tx = 32) {
  BrickGrid B;
  int tx = 8;
  while (tx < G.nx && tx < max_tx) tx <<= 1;
  B.tx = tx;
  B.ty = nthreads / tx;  // nthreads = CELLS per CTA plane
  B.bx = (G.nx + B.tx - 1) / B.tx;
  B.by = (G.ny + B.ty - 1) / B.ty;
  int lz = lz_request;
  if (lz <= 0) {
    // enough CTAs to balance 148 SMs x 2 resident CTAs, but bricks as tall as possible (less shell traffic)
    lz = 32;
    while (lz > 4 && (long long)B.bx * B.by * ((G.nzl + lz - 1) / lz) < 148 * 2 * 6) lz >>= 1;
  }
  if (lz > G.nzl) lz = G.nzl;
  if (lz < 2) lz = 2;
  B.lz = lz;
  B.bz = (G.nzl + lz - 1) / lz;
  B.ex = B.tx + 2;
  B.ey = B.ty + 2;
  B.pl = B.ex * B.ey;
  B.brick = (long long)B.pl * (lz + 2);
  return B;
}
inline size_t fused2_smem_bytes(const BrickGrid& B) { return (size_t)6 * B.pl * sizeof(double2); }
inline size_t brick_doubles2(const BrickGrid& B) { return (size_t)B.brick * B.bx * B.by * B.bz; }
inline size_t fused_smem_bytes(const BrickGrid& B) { return (size_t)6 * B.pl * sizeof(double2) + (size_t)15 * B.tx * B.ty * sizeof(double); }

// x-stage of the density scatter for one species.  p: post-collision populations of this thread's cell.
// t[g]: what arrives in this thread's column for the 9 (cy,cz) groups (own cell + left neighbour's +x movers
// + right neighbour's -x movers).  Lanes on the tile faces receive nothing from outside the tile: their adds are
// predicated off (what leaves through a face is exported separately, see ef/eg in the kernel).
// group index: 0 (0,0)  1 (+1,0)  2 (-1,0)  3 (0,+1)  4 (0,-1)  5 (+1,+1)  6 (-1,-1)  7 (+1,-1)  8 (-1,+1)
__device__ __forceinline__ double scatter_x1(double own, double to_right, double to_left, double has_left, double has_right, int width) {
  const double l = __shfl_up_sync(0xffffffffu, to_right, 1, width);
  const double r = __shfl_down_sync(0xffffffffu, to_left, 1, width);
  // has_* are exactly 1.0 or 0.0: the fma is an exact masked add (same rounding as own + l + r)
  return fma(r, has_right, fma(l, has_left, own));
}
__device__ __forceinline__ void scatter_x(const double (&p)[Q], int tx, int width, double (&t)[9]) {
  const double hl = tx != 0 ? 1. : 0., hr = tx != width - 1 ? 1. : 0.;
  t[0] = scatter_x1(p[0], p[1], p[2], hl, hr, width);
  t[1] = scatter_x1(p[3], p[7], p[10], hl, hr, width);
  t[2] = scatter_x1(p[4], p[9], p[8], hl, hr, width);
  t[3] = scatter_x1(p[5], p[15], p[18], hl, hr, width);
  t[4] = scatter_x1(p[6], p[17], p[16], hl, hr, width);
  t[5] = p[11];
  t[6] = p[12];
  t[7] = p[13];
  t[8] = p[14];
}

// 32-bit byte offsets inside one component + one (warp-uniform) 64-bit base per component: the pull addresses
// cost one integer add per direction instead of a 64-bit multiply-add per load.
__device__ __forceinline__ double ld_off(const double* __restrict__ base, unsigned byte_off) {
  return __ldg(reinterpret_cast<const double*>(reinterpret_cast<const char*>(base) + byte_off));
}
__device__ __forceinline__ void prefetch_l2(const double* __restrict__ base, unsigned byte_off) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(base) + byte_off));
}
__device__ __forceinline__ void st_off(double* __restrict__ base, unsigned byte_off, double v) {
  *reinterpret_cast<double*>(reinterpret_cast<char*>(base) + byte_off) = v;
}

template <bool NOISE, bool PREFETCH, int NT>
__global__ void __launch_bounds__(NT, 512 / NT)
k_step_fused(Geom G, BrickGrid B, DevParams P, long long step, const double* __restrict__ X, double* __restrict__ Xn,
             const double2* __restrict__ R, double2* __restrict__ E) {
  extern __shared__ double2 smem[];
  double2* Rs = smem;             // [3][ey][ex] rolling (rho,phi) planes zl-1, zl, zl+1
  double2* A = smem + 3 * B.pl;   // [3][ey][ex] rolling accumulators of next-step (rho,phi)
  double* Sg = reinterpret_cast<double*>(smem + 6 * B.pl);  // [15][256] parked moments 4..18 of species g
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * B.tx + tx;
  const int x0 = blockIdx.x * B.tx, y0 = blockIdx.y * B.ty, zb = blockIdx.z * B.lz;
  const int x = x0 + tx, y = y0 + ty;
  const bool active = x < G.nx && y < G.ny;
  const int vz = min(B.lz, G.nzl - zb);
  double2* Eb = E + (((long long)blockIdx.z * B.by + blockIdx.y) * B.bx + blockIdx.x) * B.brick;

  // stage one (rho,phi) plane zl (-1..nzl) of the tile + ring into slot s.  The in-plane source offsets are the
  // same for every plane: computed once (each thread owns at most STG entries of the extended plane).
  constexpr int STG = 2;  // pl = (tx+2)(ty+2) <= 2*NT for every tile shape used (tx in 8..32, tx*ty = NT >= 128)
  int soff[STG];
#pragma unroll
  for (int j = 0; j < STG; ++j) {
    const int idx = tid + j * NT;
    const int ey = idx / B.ex, exx = idx - ey * B.ex;
    int gx = (x0 - 1 + exx) % G.nx, gy = (y0 - 1 + ey) % G.ny;
    gx = gx < 0 ? gx + G.nx : gx;
    gy = gy < 0 ? gy + G.ny : gy;
    soff[j] = idx < B.pl ? gy * G.nx + gx : -1;
  }
  auto stage_plane = [&](int zl, int s) {
    const double2* Rp = R + (long long)(zl + 1) * G.plane;
#pragma unroll
    for (int j = 0; j < STG; ++j)
      if (soff[j] >= 0) Rs[s * B.pl + tid + j * NT] = __ldg(Rp + soff[j]);
  };
  for (int idx = tid; idx < 3 * B.pl; idx += NT) A[idx] = make_double2(0., 0.);
  stage_plane(zb - 1, 0);
  stage_plane(zb, 1);
  stage_plane(zb + 1, 2);
  __syncthreads();

  const int cell = (ty + 1) * B.ex + (tx + 1);  // this thread's cell in an extended plane
  // byte deltas to the periodic x-1 / x+1 and y-1 / y+1 neighbours (loop invariant); index 0: coordinate - 1 ... no:
  // dxv[0] = offset of x+1 minus offset of x (used when c_x = -1, source = x+1), dxv[1] = offset of x-1 minus x
  unsigned dxv[2], dyv[2], c_inpl;
  {
    const int xc = min(x, G.nx - 1), yc = min(y, G.ny - 1);
    dxv[0] = (unsigned)((xc == G.nx - 1 ? -(G.nx - 1) : 1) * 8);
    dxv[1] = (unsigned)((xc == 0 ? (G.nx - 1) : -1) * 8);
    dyv[0] = (unsigned)((yc == G.ny - 1 ? -(G.ny - 1) : 1) * G.nx * 8);
    dyv[1] = (unsigned)((yc == 0 ? (G.ny - 1) : -1) * G.nx * 8);
    c_inpl = (unsigned)(yc * G.nx + xc) * 8u;
  }
  const unsigned pl8 = (unsigned)G.plane * 8u;
  for (int k = 0; k < vz; ++k) {
    const int zl = zb + k;
    // slot of plane zl + d : (k + 1 + d) % 3
    const int s_m = k % 3, s_0 = (k + 1) % 3, s_p = (k + 2) % 3;
    double tf[9], tg[9];
    double ef[5], eg[5];  // edge exports (left face on lane 0, right face on lane tx-1)
    const bool left = tx == 0;
    {
      double mf[Q], mg[Q];
      unsigned c = 0;  // byte offset of this cell inside a component
      CollideCtx C;
      NoiseKey nk;
      if (active) {
        // gradients first (LBM_binary.H:134-150): only 6 doubles stay live across the population loads
        double grho[3], gphi[3];
        {
          double nr[Q], np[Q];
          nr[0] = np[0] = 0.;
          const int sl[3] = {s_m, s_0, s_p};
#pragma unroll
          for (int i = 1; i < Q; ++i) {
            const double2 v = Rs[sl[1 + cz(i)] * B.pl + cell + cy(i) * B.ex + cx(i)];
            nr[i] = v.x;
            np[i] = v.y;
          }
          gradient19(nr, grho);
          gradient19(np, gphi);
        }
        // byte offsets of the 19 pull sources (x - c_i) inside a component; < 4 GiB is checked at creation
        unsigned off[Q];
        {
          c = (unsigned)(zl + 1) * pl8 + c_inpl;
          const unsigned dz[3] = {0u - pl8, 0u, pl8};
#pragma unroll
          for (int i = 0; i < Q; ++i) {  // source = cell - c_i ; unsigned arithmetic wraps consistently
            unsigned o = c;
            if (cz(i) != 0) o += dz[1 - cz(i)];
            if (cy(i) != 0) o += dyv[(1 + cy(i)) >> 1];
            if (cx(i) != 0) o += dxv[(1 + cx(i)) >> 1];
            off[i] = o;
          }
        }
        if (PREFETCH && (tx & 15) == 0 && k + 1 < vz) {
          // next plane of this column into L2 (one request per 128 B line) while this plane computes
          const unsigned nxt = c + (unsigned)G.plane * 8u;
#pragma unroll
          for (int i = 0; i < 2 * Q; ++i) prefetch_l2(X + (long long)i * G.comp, nxt);
        }
        {
          double f[Q];
          // species g first: its non-conserved moments wait in shared memory while species f is processed
#pragma unroll
          for (int i = 0; i < Q; ++i) f[i] = ld_off(X + (long long)(Q + i) * G.comp, off[i]);
          moments(f, mg);
#pragma unroll
          for (int a = 4; a < Q; ++a) Sg[(a - 4) * NT + tid] = mg[a];
#pragma unroll
          for (int i = 0; i < Q; ++i) f[i] = ld_off(X + (long long)i * G.comp, off[i]);
          moments(f, mf);
        }
        nk = make_noise_key(P.keys, (unsigned long long)cell_global(G, x, y, zl), step);
        collide_prepare<NOISE>(P, grho, gphi, nk, mf, mg, C);
        collide_species<NOISE, 0>(P, nk, C, mf);
      } else {
#pragma unroll
        for (int i = 0; i < Q; ++i) mf[i] = 0.;
      }
      double p[Q];
      populations(mf, p);
      if (active) {
#pragma unroll
        for (int i = 0; i < Q; ++i) st_off(Xn + (long long)i * G.comp, c, p[i]);
      }
      scatter_x(p, tx, B.tx, tf);
      // what leaves through the left face (lane 0) / right face (lane tx-1); unused on the other lanes
      ef[0] = left ? p[2] : p[1]; ef[1] = left ? p[10] : p[7]; ef[2] = left ? p[8] : p[9];
      ef[3] = left ? p[18] : p[15]; ef[4] = left ? p[16] : p[17];
      if (active) {
#pragma unroll
        for (int a = 4; a < Q; ++a) mg[a] = Sg[(a - 4) * NT + tid];
        collide_species<NOISE, 1>(P, nk, C, mg);
      } else {
#pragma unroll
        for (int i = 0; i < Q; ++i) mg[i] = 0.;
      }
      populations(mg, p);
      if (active) {
#pragma unroll
        for (int i = 0; i < Q; ++i) st_off(Xn + (long long)(Q + i) * G.comp, c, p[i]);
      }
      scatter_x(p, tx, B.tx, tg);
      eg[0] = left ? p[2] : p[1]; eg[1] = left ? p[10] : p[7]; eg[2] = left ? p[8] : p[9];
      eg[3] = left ? p[18] : p[15]; eg[4] = left ? p[16] : p[17];
    }
    // a tile narrower than two lanes would need both faces on one lane; tx >= 8 always
    const bool edge = (tx == 0) || (tx == B.tx - 1);
    const int ecell = (ty + 1) * B.ex + (tx == 0 ? 0 : B.tx + 1);

    __syncthreads();  // S0: everyone is done reading Rs slot s_m (plane zl-1) and the last write-out is finished
    if (zl + 2 <= G.nzl) stage_plane(zl + 2, s_m);
    auto add = [&](int slot, int at, double a, double b) {
      double2 v = A[slot * B.pl + at];
      v.x += a;
      v.y += b;
      A[slot * B.pl + at] = v;
    };
    // phase cy = 0 : own row.  groups 0 (cz 0), 3 (cz +1), 4 (cz -1)
    add(s_0, cell, tf[0], tg[0]);
    add(s_p, cell, tf[3], tg[3]);
    add(s_m, cell, tf[4], tg[4]);
    if (edge) {
      add(s_0, ecell, ef[0], eg[0]);
      add(s_p, ecell, ef[3], eg[3]);
      add(s_m, ecell, ef[4], eg[4]);
    }
    __syncthreads();
    // phase cy = +1 : row above.  groups 1 (cz 0), 5 (cz +1), 7 (cz -1)
    add(s_0, cell + B.ex, tf[1], tg[1]);
    add(s_p, cell + B.ex, tf[5], tg[5]);
    add(s_m, cell + B.ex, tf[7], tg[7]);
    if (edge) add(s_0, ecell + B.ex, ef[1], eg[1]);
    __syncthreads();
    // phase cy = -1 : row below.  groups 2 (cz 0), 8 (cz +1), 6 (cz -1)
    add(s_0, cell - B.ex, tf[2], tg[2]);
    add(s_p, cell - B.ex, tf[8], tg[8]);
    add(s_m, cell - B.ex, tf[6], tg[6]);
    if (edge) add(s_0, ecell - B.ex, ef[2], eg[2]);
    __syncthreads();
    // plane zl-1 has received everything this brick can give it: write it out (extended plane k) and recycle
    for (int idx = tid; idx < B.pl; idx += NT) {
      Eb[(long long)k * B.pl + idx] = A[s_m * B.pl + idx];
      A[s_m * B.pl + idx] = make_double2(0., 0.);
    }
  }
  __syncthreads();
  // the two planes still in flight: zl = zb+vz-1 (extended plane vz) and the top shell (vz+1)
  for (int idx = tid; idx < B.pl; idx += NT) {
    Eb[(long long)vz * B.pl + idx] = A[(vz % 3) * B.pl + idx];
    Eb[(long long)(vz + 1) * B.pl + idx] = A[((vz + 1) % 3) * B.pl + idx];
  }
}

// ---------------------------------------------------------------------------------------------
// Species-split variant of the fused step: TWO threads per cell, lane l (l < 16) handles species f and lane
// l + 16 species g of the same cell (a warp = 16 consecutive cells x 2 species).  Each thread carries 19
// instead of 38 populations, so the kernel fits 3 CTAs (24 warps) per SM instead of 16 warps, and the
// per-warp critical path between barriers is half as long: the collision is latency/issue bound
// (profiles/r1_*), not DRAM bound, and more resident warps is what it needs.
// The code is species-uniform: the species enters only as data (base pointers, rate, sign of the momentum
// noise, Philox block offset), so the two half-warps never diverge.  What each thread needs from the other
// species (density, momentum, acceleration, real velocity) crosses with shfl.xor 16.

// everything the relaxation of species s needs: own real velocity u, barycentric velocity vb, momentum noise xi
template <bool NOISE>
__device__ __forceinline__ void pair_hydro(const DevParams& P, int s, double dens_s, const double (&j_s)[3], const double (&a_s)[3],
                                           const float (&n3)[3], double (&u_s)[3], double (&vb)[3], double (&xi_s)[3]) {
  const unsigned full = 0xffffffffu;
  const double dens_o = __shfl_xor_sync(full, dens_s, 16);
  const bool has_s = fabs(dens_s) > (double)FLT_EPSILON;
  const double inv_s = has_s ? 1. / dens_s : 0.;
  const double inv_o = __shfl_xor_sync(full, inv_s, 16);
  const double inv_tot = 1. / (dens_s + dens_o);  // unguarded like the reference; a+b is the same in both lanes
  const double fric_s = s ? P.fric_g : P.fric_f;
  double amp = 0.;
  if (NOISE) amp = sqrt(P.amp_j * fabs(dens_s * dens_o * inv_tot));
  const double sign = s ? -1. : 1.;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double j_o = __shfl_xor_sync(full, j_s[k], 16), a_o = __shfl_xor_sync(full, a_s[k], 16);
    const double ub_s = j_s[k] * inv_s, ub_o = j_o * inv_o;
    xi_s[k] = NOISE ? sign * (amp * (double)n3[k]) : 0.;
    const double d = (ub_s - ub_o) + 0.5 * (a_s[k] - a_o);
    u_s[k] = ub_s + 0.5 * a_s[k] - fric_s * dens_o * inv_tot * d + 0.5 * (xi_s[k] * inv_s);
    const double u_o = __shfl_xor_sync(full, u_s[k], 16);
    // (rho u_f + phi u_g)/(rho+phi), LBM_binary.H:471; no fma contraction so that both lanes get the same bits
    vb[k] = __dmul_rn(__dadd_rn(__dmul_rn(dens_s, u_s[k]), __dmul_rn(dens_o, u_o)), inv_tot);
  }
}

template <bool NOISE>
__global__ void __launch_bounds__(256, 3)
k_step_fused2(Geom G, BrickGrid B, DevParams P, long long step, const double* __restrict__ X, double* __restrict__ Xn,
              const double2* __restrict__ R, double2* __restrict__ E) {
  constexpr int NT = 256;
  extern __shared__ double2 smem[];
  double2* Rs = smem;             // [3][ey][ex] rolling (rho,phi) planes zl-1, zl, zl+1
  double2* A = smem + 3 * B.pl;   // [3][ey][ex] rolling accumulators of next-step (rho,phi)
  double* Ad = reinterpret_cast<double*>(A);
  const double* Rd = reinterpret_cast<const double*>(Rs);
  const int tid = threadIdx.x, lane = tid & 31, s = lane >> 4;
  const int cidx = (tid >> 5) * 16 + (lane & 15);  // cell of the tile (tx*ty = 128 cells)
  const int tx = cidx % B.tx, ty = cidx / B.tx;
  const int x0 = blockIdx.x * B.tx, y0 = blockIdx.y * B.ty, zb = blockIdx.z * B.lz;
  const int x = x0 + tx, y = y0 + ty;
  const bool active = x < G.nx && y < G.ny;
  const int vz = min(B.lz, G.nzl - zb);
  double2* Eb = E + (((long long)blockIdx.z * B.by + blockIdx.y) * B.bx + blockIdx.x) * B.brick;
  const double* __restrict__ Xs = X + (long long)(s * Q) * G.comp;
  double* __restrict__ Xns = Xn + (long long)(s * Q) * G.comp;

  auto stage_plane = [&](int zl, int slot) {
    const double2* Rp = R + (long long)(zl + 1) * G.plane;
    for (int idx = tid; idx < B.pl; idx += NT) {
      const int ey = idx / B.ex, exx = idx - ey * B.ex;
      int gx = (x0 - 1 + exx) % G.nx, gy = (y0 - 1 + ey) % G.ny;
      gx = gx < 0 ? gx + G.nx : gx;
      gy = gy < 0 ? gy + G.ny : gy;
      Rs[slot * B.pl + idx] = __ldg(Rp + (long long)gy * G.nx + gx);
    }
  };
  for (int idx = tid; idx < 3 * B.pl; idx += NT) A[idx] = make_double2(0., 0.);
  stage_plane(zb - 1, 0);
  stage_plane(zb, 1);
  stage_plane(zb + 1, 2);
  __syncthreads();

  // in-component byte offsets of the in-plane parts of the 19 pull sources (plane part added per plane)
  const unsigned xs[3] = {(unsigned)(x == 0 ? G.nx - 1 : x - 1), (unsigned)min(x, G.nx - 1), (unsigned)(x >= G.nx - 1 ? 0 : x + 1)};
  const int yc = min(y, G.ny - 1);
  const unsigned yr[3] = {(unsigned)(yc == 0 ? G.ny - 1 : yc - 1) * (unsigned)G.nx, (unsigned)yc * (unsigned)G.nx,
                          (unsigned)(yc == G.ny - 1 ? 0 : yc + 1) * (unsigned)G.nx};
  const int cell = (ty + 1) * B.ex + (tx + 1);  // this thread's cell in an extended plane
  const bool left = tx == 0, edge = left || tx == B.tx - 1;
  const int ecell = (ty + 1) * B.ex + (left ? 0 : B.tx + 1);
  const double rate = s ? P.rate_g : P.rate_f;

  for (int k = 0; k < vz; ++k) {
    const int zl = zb + k;
    const int s_m = k % 3, s_0 = (k + 1) % 3, s_p = (k + 2) % 3;  // slot of plane zl + d : (k + 1 + d) % 3
    double t[9], e[5];
    {
      // acceleration of this species from the gradient of the OTHER species' density (LBM_binary.H:134-150, 254-255)
      double go[3];
      {
        double n[Q];
        n[0] = 0.;
        const int sl[3] = {s_m, s_0, s_p};
#pragma unroll
        for (int i = 1; i < Q; ++i) n[i] = Rd[2 * (sl[1 + cz(i)] * B.pl + cell + cy(i) * B.ex + cx(i)) + (1 - s)];
        gradient19(n, go);
      }
      const unsigned pl8 = (unsigned)G.plane * 8u;
      const unsigned zp[3] = {(unsigned)zl * pl8, (unsigned)(zl + 1) * pl8, (unsigned)(zl + 2) * pl8};
      const unsigned c = zp[1] + (yr[1] + xs[1]) * 8u;
      double m[Q];
      {
        double f[Q];
#pragma unroll
        for (int i = 0; i < Q; ++i) {
          const unsigned off = zp[1 - cz(i)] + (yr[1 - cy(i)] + xs[1 - cx(i)]) * 8u;
          f[i] = active ? ld_off(Xs + (long long)i * G.comp, off) : 0.;
        }
        moments(f, m);
      }
      const bool has_s = fabs(m[0]) > (double)FLT_EPSILON;
      double a_s[3], u_s[3], vb[3], xi_s[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) a_s[d] = has_s ? P.acc_coef * go[d] : 0.;
      NoiseKey nk;
      float n0[4] = {0.f, 0.f, 0.f, 0.f};
      if (NOISE) {
        nk = make_noise_key(P.keys, (unsigned long long)cell_global(G, x, y, zl), step);
        normals4(nk, 0, n0);
      }
      {
        const double j_s[3] = {m[1], m[2], m[3]};
        const float n3[3] = {n0[0], n0[1], n0[2]};
        pair_hydro<NOISE>(P, s, m[0], j_s, a_s, n3, u_s, vb, xi_s);
      }
      relax_species(rate, P.force_pf, m[0], vb, u_s, a_s, m);
      if (NOISE) {
#pragma unroll
        for (int d = 0; d < 3; ++d) m[1 + d] += xi_s[d];
        const double sa = sqrt(P.amp_s * fabs(m[0]));
        float nb[4];
#pragma unroll
        for (int a = 4; a < Q; ++a) {
          if (((a - 4) & 3) == 0) normals4(nk, 1 + 4 * s + ((a - 4) >> 2), nb);  // = mode_index(s, a) >> 2
          m[a] += (sqrt_bnorm(a) * sa) * (double)nb[(a - 4) & 3];
        }
      }
      double p[Q];
      populations(m, p);
      if (active) {
#pragma unroll
        for (int i = 0; i < Q; ++i) st_off(Xns + (long long)i * G.comp, c, p[i]);
      } else {
#pragma unroll
        for (int i = 0; i < Q; ++i) p[i] = 0.;
      }
      scatter_x(p, tx, B.tx, t);
      // what leaves through the left face (lane 0) / right face (lane tx-1); unused on the other lanes
      e[0] = left ? p[2] : p[1]; e[1] = left ? p[10] : p[7]; e[2] = left ? p[8] : p[9];
      e[3] = left ? p[18] : p[15]; e[4] = left ? p[16] : p[17];
    }

    __syncthreads();  // S0: everyone is done reading Rs slot s_m (plane zl-1) and the last write-out is finished
    if (zl + 2 <= G.nzl) stage_plane(zl + 2, s_m);
    auto add = [&](int slot, int at, double v) { Ad[2 * (slot * B.pl + at) + s] += v; };
    // phase cy = 0 : own row.  groups 0 (cz 0), 3 (cz +1), 4 (cz -1)
    add(s_0, cell, t[0]);
    add(s_p, cell, t[3]);
    add(s_m, cell, t[4]);
    if (edge) {
      add(s_0, ecell, e[0]);
      add(s_p, ecell, e[3]);
      add(s_m, ecell, e[4]);
    }
    __syncthreads();
    // phase cy = +1 : row above.  groups 1 (cz 0), 5 (cz +1), 7 (cz -1)
    add(s_0, cell + B.ex, t[1]);
    add(s_p, cell + B.ex, t[5]);
    add(s_m, cell + B.ex, t[7]);
    if (edge) add(s_0, ecell + B.ex, e[1]);
    __syncthreads();
    // phase cy = -1 : row below.  groups 2 (cz 0), 8 (cz +1), 6 (cz -1)
    add(s_0, cell - B.ex, t[2]);
    add(s_p, cell - B.ex, t[8]);
    add(s_m, cell - B.ex, t[6]);
    if (edge) add(s_0, ecell - B.ex, e[2]);
    __syncthreads();
    // plane zl-1 has received everything this brick can give it: write it out (extended plane k) and recycle
    for (int idx = tid; idx < B.pl; idx += NT) {
      Eb[(long long)k * B.pl + idx] = A[s_m * B.pl + idx];
      A[s_m * B.pl + idx] = make_double2(0., 0.);
    }
  }
  __syncthreads();
  for (int idx = tid; idx < B.pl; idx += NT) {
    Eb[(long long)vz * B.pl + idx] = A[(vz % 3) * B.pl + idx];
    Eb[(long long)(vz + 1) * B.pl + idx] = A[((vz + 1) % 3) * B.pl + idx];
  }
}

// (rho,phi)(x,y,zl) = sum over the bricks whose extended box contains the cell, fixed order
// (dz: 0,-1,+1; dy: 0,-1,+1; dx: 0,-1,+1), grouped per dz so that the part a neighbouring slab contributes
// (dz = -1 at the bottom plane, +1 at the top plane) is ONE addend: bit-identical for any slab count.
// zl = -1 and zl = nzl give this slab's contribution to the neighbour's boundary plane.
__global__ void __launch_bounds__(256) k_fold(Geom G, BrickGrid B, const double2* __restrict__ E, double2* __restrict__ R) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, zl = (int)blockIdx.z - 1;
  if (x >= G.nx || y >= G.ny) return;
  int nbx[3], ncx[3], nby[3], ncy[3], nbz[3], ncz[3];  // candidate (brick, extended coordinate) per axis; brick -1 = none
  {
    const int b = x / B.tx, l = x - b * B.tx, v = min(B.tx, G.nx - b * B.tx);
    const int bm = (b + B.bx - 1) % B.bx, bp = (b + 1) % B.bx;
    nbx[0] = b; ncx[0] = l + 1;
    nbx[1] = (l == 0) ? bm : -1;     ncx[1] = min(B.tx, G.nx - bm * B.tx) + 1;
    nbx[2] = (l == v - 1) ? bp : -1; ncx[2] = 0;
  }
  {
    const int b = y / B.ty, l = y - b * B.ty, v = min(B.ty, G.ny - b * B.ty);
    const int bm = (b + B.by - 1) % B.by, bp = (b + 1) % B.by;
    nby[0] = b; ncy[0] = l + 1;
    nby[1] = (l == 0) ? bm : -1;     ncy[1] = min(B.ty, G.ny - bm * B.ty) + 1;
    nby[2] = (l == v - 1) ? bp : -1; ncy[2] = 0;
  }
  if (zl < 0) {
    nbz[0] = -1; ncz[0] = 0; nbz[1] = -1; ncz[1] = 0; nbz[2] = 0; ncz[2] = 0;
  } else if (zl >= G.nzl) {
    nbz[0] = -1; ncz[0] = 0; nbz[2] = -1; ncz[2] = 0;
    nbz[1] = B.bz - 1; ncz[1] = min(B.lz, G.nzl - (B.bz - 1) * B.lz) + 1;
  } else {
    const int b = zl / B.lz, l = zl - b * B.lz, v = min(B.lz, G.nzl - b * B.lz);
    nbz[0] = b; ncz[0] = l + 1;
    nbz[1] = (l == 0 && b > 0) ? b - 1 : -1;            ncz[1] = B.lz + 1;  // a lower brick is always full height
    nbz[2] = (l == v - 1 && b + 1 < B.bz) ? b + 1 : -1; ncz[2] = 0;
  }
  double2 tot = make_double2(0., 0.);
  bool tot_set = false;
#pragma unroll
  for (int dz = 0; dz < 3; ++dz) {
    if (nbz[dz] < 0) continue;
    double2 s = make_double2(0., 0.);
    bool s_set = false;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      if (nby[dy] < 0) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        if (nbx[dx] < 0) continue;
        const long long brick = ((long long)nbz[dz] * B.by + nby[dy]) * B.bx + nbx[dx];
        const double2 v = __ldg(E + brick * B.brick + (long long)ncz[dz] * B.pl + ncy[dy] * B.ex + ncx[dx]);
        if (s_set) { s.x += v.x; s.y += v.y; } else { s = v; s_set = true; }
      }
    }
    if (tot_set) { tot.x += s.x; tot.y += s.y; } else { tot = s; tot_set = true; }
  }
  R[(long long)(zl + 1) * G.plane + (long long)y * G.nx + x] = tot;
}

// same local sums straight from the populations (after a restart upload, before the first fused step):
// sum_i X_i(x - c_i) over the source planes this slab owns, i = 0..18 in order.
__global__ void __launch_bounds__(256) k_density_partial(Geom G, const double* __restrict__ X, double2* __restrict__ R) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, zl = (int)blockIdx.z - 1;
  if (x >= G.nx || y >= G.ny) return;
  CellIdx I = cell_idx(G, x, y, zl < 0 ? 0 : (zl >= G.nzl ? G.nzl - 1 : zl));
  // rebuild the plane offsets for the true zl (cell_idx clamps nothing; ghost rows are addressable)
  I.zpl[0] = (long long)zl * G.plane;
  I.zpl[1] = (long long)(zl + 1) * G.plane;
  I.zpl[2] = (long long)(zl + 2) * G.plane;
  double rho = 0., phi = 0.;
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    const int zs = zl - cz(i);  // source plane
    if (zs < 0 || zs >= G.nzl) continue;
    const long long a = nbr<-1>(I, i);
    rho += __ldg(X + (long long)i * G.comp + a);
    phi += __ldg(X + (long long)(Q + i) * G.comp + a);
  }
  R[I.zpl[1] + I.yrow[1] + x] = make_double2(rho, phi);
}

// receiving side of the density part of a halo message:
//   boundary plane:  R <- local + Ez(neighbour's contribution)      ghost plane:  R <- Pz(neighbour's local) + mine
__global__ void k_merge_density_halo(long long plane, const double2* __restrict__ Pz, const double2* __restrict__ Ez,
                                     double2* __restrict__ Rb, double2* __restrict__ Rg) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= plane) return;
  const double2 p = Pz[i], e = Ez[i];
  double2 b = Rb[i], g = Rg[i];
  b.x += e.x; b.y += e.y;
  g.x = p.x + g.x; g.y = p.y + g.y;
  Rb[i] = b;
  Rg[i] = g;
}

}  // namespace bflbm
