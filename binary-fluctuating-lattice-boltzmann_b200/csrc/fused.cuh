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

inline BrickGrid make_brick_grid(const Geom& G, int lz_request = 0) {
  BrickGrid B;
  int tx = 8;
  while (tx < G.nx && tx < 32) tx <<= 1;
  B.tx = tx;
  B.ty = 256 / tx;
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
inline size_t brick_doubles2(const BrickGrid& B) { return (size_t)B.brick * B.bx * B.by * B.bz; }
inline size_t fused_smem_bytes(const BrickGrid& B) { return (size_t)6 * B.pl * sizeof(double2); }

__device__ __forceinline__ double shfl_from_left(double v, int tx, int width) {
  const double r = __shfl_up_sync(0xffffffffu, v, 1, width);
  return tx == 0 ? 0. : r;
}
__device__ __forceinline__ double shfl_from_right(double v, int tx, int width) {
  const double r = __shfl_down_sync(0xffffffffu, v, 1, width);
  return tx == width - 1 ? 0. : r;
}

// x-stage of the density scatter for one species.  p: post-collision populations of this thread's cell.
// t[g]: what arrives in this thread's column for the 9 (cy,cz) groups; sm[5]/sp[5]: what leaves through the
// left/right face of the tile (meaningful on lanes tx==0 / tx==width-1), for the 5 groups that have cx != 0.
// group index: 0 (0,0)  1 (+1,0)  2 (-1,0)  3 (0,+1)  4 (0,-1)  5 (+1,+1)  6 (-1,-1)  7 (+1,-1)  8 (-1,+1)
__device__ __forceinline__ void scatter_x(const double (&p)[Q], int tx, int width, double (&t)[9]) {
  t[0] = p[0] + shfl_from_left(p[1], tx, width) + shfl_from_right(p[2], tx, width);
  t[1] = p[3] + shfl_from_left(p[7], tx, width) + shfl_from_right(p[10], tx, width);
  t[2] = p[4] + shfl_from_left(p[9], tx, width) + shfl_from_right(p[8], tx, width);
  t[3] = p[5] + shfl_from_left(p[15], tx, width) + shfl_from_right(p[18], tx, width);
  t[4] = p[6] + shfl_from_left(p[17], tx, width) + shfl_from_right(p[16], tx, width);
  t[5] = p[11];
  t[6] = p[12];
  t[7] = p[13];
  t[8] = p[14];
}

template <bool NOISE>
__global__ void __launch_bounds__(256, 2)
k_step_fused(Geom G, BrickGrid B, DevParams P, long long step, const double* __restrict__ X, double* __restrict__ Xn,
             const double2* __restrict__ R, double2* __restrict__ E) {
  extern __shared__ double2 smem[];
  double2* Rs = smem;             // [3][ey][ex] rolling (rho,phi) planes zl-1, zl, zl+1
  double2* A = smem + 3 * B.pl;   // [3][ey][ex] rolling accumulators of next-step (rho,phi)
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * B.tx + tx;
  const int x0 = blockIdx.x * B.tx, y0 = blockIdx.y * B.ty, zb = blockIdx.z * B.lz;
  const int x = x0 + tx, y = y0 + ty;
  const bool active = x < G.nx && y < G.ny;
  const int vz = min(B.lz, G.nzl - zb);
  double2* Eb = E + (((long long)blockIdx.z * B.by + blockIdx.y) * B.bx + blockIdx.x) * B.brick;

  // stage one (rho,phi) plane zl (-1..nzl) of the tile + ring into slot s
  auto stage_plane = [&](int zl, int s) {
    const double2* Rp = R + (long long)(zl + 1) * G.plane;
    for (int idx = tid; idx < B.pl; idx += 256) {
      const int ey = idx / B.ex, exx = idx - ey * B.ex;
      int gx = (x0 - 1 + exx) % G.nx, gy = (y0 - 1 + ey) % G.ny;
      gx = gx < 0 ? gx + G.nx : gx;
      gy = gy < 0 ? gy + G.ny : gy;
      Rs[s * B.pl + idx] = __ldg(Rp + (long long)gy * G.nx + gx);
    }
  };
  for (int idx = tid; idx < 3 * B.pl; idx += 256) A[idx] = make_double2(0., 0.);
  stage_plane(zb - 1, 0);
  stage_plane(zb, 1);
  stage_plane(zb + 1, 2);
  __syncthreads();

  const int cell = (ty + 1) * B.ex + (tx + 1);  // this thread's cell in an extended plane
  for (int k = 0; k < vz; ++k) {
    const int zl = zb + k;
    // slot of plane zl + d : (k + 1 + d) % 3
    const int s_m = k % 3, s_0 = (k + 1) % 3, s_p = (k + 2) % 3;
    double tf[9], tg[9];
    double ef[5], eg[5];  // edge exports (left face on lane 0, right face on lane tx-1)
#pragma unroll
    for (int j = 0; j < 5; ++j) ef[j] = eg[j] = 0.;
    {
      double mf[Q], mg[Q];
      long long c = 0;
      if (active) {
        const CellIdx I = cell_idx(G, x, y, zl);
        c = I.zpl[1] + I.yrow[1] + x;
        {
          double f[Q];
          pull19(X, G, I, f);
          moments(f, mf);
          pull19(X + (long long)Q * G.comp, G, I, f);
          moments(f, mg);
        }
        // gradients from the staged neighbourhood (LBM_binary.H:134-150)
        double nr[Q], np[Q], grho[3], gphi[3];
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
        const NoiseKey nk = make_noise_key(P.seed, (unsigned long long)cell_global(G, x, y, zl), step);
        collide_cell<NOISE>(P, grho, gphi, nk, mf, mg);
      } else {
#pragma unroll
        for (int i = 0; i < Q; ++i) mf[i] = mg[i] = 0.;
      }
      double p[Q];
      populations(mf, p);
      if (active) {
#pragma unroll
        for (int i = 0; i < Q; ++i) Xn[(long long)i * G.comp + c] = p[i];
      }
      scatter_x(p, tx, B.tx, tf);
      if (tx == 0)        { ef[0] = p[2]; ef[1] = p[10]; ef[2] = p[8]; ef[3] = p[18]; ef[4] = p[16]; }
      if (tx == B.tx - 1) { ef[0] = p[1]; ef[1] = p[7];  ef[2] = p[9]; ef[3] = p[15]; ef[4] = p[17]; }
      populations(mg, p);
      if (active) {
#pragma unroll
        for (int i = 0; i < Q; ++i) Xn[(long long)(Q + i) * G.comp + c] = p[i];
      }
      scatter_x(p, tx, B.tx, tg);
      if (tx == 0)        { eg[0] = p[2]; eg[1] = p[10]; eg[2] = p[8]; eg[3] = p[18]; eg[4] = p[16]; }
      if (tx == B.tx - 1) { eg[0] = p[1]; eg[1] = p[7];  eg[2] = p[9]; eg[3] = p[15]; eg[4] = p[17]; }
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
    for (int idx = tid; idx < B.pl; idx += 256) {
      Eb[(long long)k * B.pl + idx] = A[s_m * B.pl + idx];
      A[s_m * B.pl + idx] = make_double2(0., 0.);
    }
  }
  __syncthreads();
  // the two planes still in flight: zl = zb+vz-1 (extended plane vz) and the top shell (vz+1)
  for (int idx = tid; idx < B.pl; idx += 256) {
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
