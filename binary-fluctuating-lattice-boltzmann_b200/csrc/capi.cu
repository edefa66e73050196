// C ABI of the B200-native fluctuating binary D3Q19 step (see include/bflbm.h for the contract and the
// reference interface each entry point replaces).  Host-side orchestration only; all arithmetic is in
// kernels.cuh / fused.cuh.  There is deliberately no CPU path: without a CUDA device every call fails.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "../../include/bflbm.h"
#include "fused.cuh"
#include "kernels.cuh"

using namespace bflbm;

namespace {
thread_local std::string g_err;
int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CU(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(BFLBM_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define CHECK_H(h)                                        \
  if (!(h)) return fail(BFLBM_ERR_ARG, "null lattice handle")
}  // namespace

struct bflbm_lattice {
  Geom G{};
  bflbm_params prm{};
  DevParams dp{};
  int device = 0;
  bool whole_box = true;
  bool initialized = false;
  int algo = 0;  // 0 fused one-pass, 1 two-pass
  int lz_request = 0;
  int cta_threads = 256;  // threads per CTA of the fused kernel (BFLBM_CTA_THREADS=128|256)
  bool rate1_fast_path = true;  // use the rate == 1 specialisation when tau_f = tau_g = 1/2 (BFLBM_RATE1=0 disables it)
  long long step = 0;
  long long launches = 0;
  size_t bytes = 0;

  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // slab step with the halo exchange overlapped: the first and last brick row, the slab-face fold and the packing run
  // on `stream` (where the caller then queues its NCCL send/recv); the interior rows run on `aux` meanwhile
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_prev = nullptr, ev_ends = nullptr, ev_interior = nullptr;
  bool overlap = true;          // BFLBM_OVERLAP=0: everything on `stream`, exchange after the whole step kernel
  bool interior_pending = false;

  double* X[2] = {nullptr, nullptr};
  int cur = 0;
  double2* R = nullptr;

  // fused algorithm
  BrickGrid B{};
  double2* E[2] = {nullptr, nullptr};  // extended boxes of the last two steps (ping-pong): E[ecur] is the newest
  int ecur = 0;
  bool e_valid = false;        // E[ecur] holds the densities of the current state (the step kernel may fold from it)
  bool r_stale = false;        // R is current only on brick-face planes; k_fold<2>(E[ecur]) completes it on demand
  bool fold_in_staging = true; // BFLBM_FOLD_IN_STAGING=0: separate full fold pass every step (first design)

  // halo messages (slab) : [side] ; layout see pack_halo()
  double* send[2] = {nullptr, nullptr};
  double* recv[2] = {nullptr, nullptr};
  size_t halo_doubles = 0;

  // staging for host transfers and diagnostics
  double* stage = nullptr;
  size_t stage_doubles = 0;
  double* diag_partial = nullptr;
  unsigned long long* diag_count = nullptr;
  size_t diag_blocks = 0;

  dim3 block, grid_xy;  // thread-per-cell kernels: grid = (grid_xy.x, grid_xy.y, planes)

  // optional per-kernel timing: events ev[0..4] bracket {step kernel, fold, pack, unpack}
  bool profiling = false;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  double prof_ms[4] = {0., 0., 0., 0.};
  long long prof_steps = 0;
};

namespace {

int ceil_div(int a, int b) { return (a + b - 1) / b; }

// the eight variants of the step kernel for one CTA size; F = kernel functor applied to each instantiation
template <int NT, class Fn>
cudaError_t for_each_fused(Fn fn) {
  cudaError_t e = cudaSuccess;
#define BFLBM_EACH(N, R1, FU) if (e == cudaSuccess) e = fn((const void*)k_step_fused<N, R1, FU, NT>, R1)
  BFLBM_EACH(false, false, false); BFLBM_EACH(false, false, true); BFLBM_EACH(false, true, false); BFLBM_EACH(false, true, true);
  BFLBM_EACH(true, false, false);  BFLBM_EACH(true, false, true);  BFLBM_EACH(true, true, false);  BFLBM_EACH(true, true, true);
#undef BFLBM_EACH
  return e;
}
// Dynamic shared memory of the fused kernels.  The attribute is per FUNCTION and process-wide, not per lattice, so it is
// set once to the largest need over every brick shape make_brick_grid can produce (tx = 8, 16, 32): lattices of different
// shape or tau can then coexist.
template <int NT>
size_t fused_smem_max(bool rate1) {
  size_t m = 0;
  for (int tx = 8; tx <= 32; tx <<= 1) {
    BrickGrid B{};
    B.tx = tx; B.ty = NT / tx; B.ex = B.tx + 2; B.ey = B.ty + 2; B.pl = B.ex * B.ey;
    m = std::max(m, fused_smem_bytes(B, rate1));
  }
  return m;
}
cudaError_t set_fused_smem() {
  static bool done[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || (dev < 64 && done[dev])) return e;
  e = for_each_fused<128>([&](const void* k, bool rate1) {
    return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem_max<128>(rate1));
  });
  if (e == cudaSuccess)
    e = for_each_fused<256>([&](const void* k, bool rate1) {
      return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem_max<256>(rate1));
    });
  // BFLBM_CARVEOUT=<percent of the SM's shared memory>: experiment knob (the driver otherwise sizes the carve-out to the
  // resident CTAs' need; what is left of the 256 KB is L1)
  if (const char* cv = getenv("BFLBM_CARVEOUT")) {
    const int pct = atoi(cv);
    auto setc = [&](const void* k, bool) { return cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct); };
    if (e == cudaSuccess) e = for_each_fused<128>(setc);
    if (e == cudaSuccess) e = for_each_fused<256>(setc);
  }
  if (e == cudaSuccess && dev < 64) done[dev] = true;
  return e;
}

inline void mark(bflbm_lattice* h, int i) {
  if (h->profiling) cudaEventRecord(h->ev[i], h->stream);
}
// call after mark(4): folds the four intervals of one step into the totals (synchronises the stream)
inline void profile_collect(bflbm_lattice* h) {
  if (!h->profiling) return;
  cudaEventSynchronize(h->ev[4]);
  for (int i = 0; i < 4; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]) == cudaSuccess) h->prof_ms[i] += ms;
  }
  ++h->prof_steps;
}

int set_device(const bflbm_lattice* h) {
  CU(cudaSetDevice(h->device));
  return 0;
}

int validate_params(const bflbm_params* p) {
  if (!p) return fail(BFLBM_ERR_ARG, "null params");
  if (!(p->tau_f > 0.) || !(p->tau_g > 0.)) return fail(BFLBM_ERR_ARG, "tau_f and tau_g must be > 0");
  if (p->alpha1 != 0.) return fail(BFLBM_ERR_ARG, "alpha1 must be 0: it does not enter the reference dynamics (LBM_binary.H:256-257)");
  if (p->kBT < 0.) return fail(BFLBM_ERR_ARG, "kBT must be >= 0");
  if (!(p->kappa > 0.)) return fail(BFLBM_ERR_ARG, "kappa must be > 0");
  return 0;
}

void derive(bflbm_lattice* h) {
  const bflbm_params& p = h->prm;
  DevParams& d = h->dp;
  d.rate_f = 1. / (p.tau_f * (1. + 0.5 / p.tau_f));
  d.rate_g = 1. / (p.tau_g * (1. + 0.5 / p.tau_g));
  d.fric_f = 0.5 / (p.tau_f + 0.5);
  d.fric_g = 0.5 / (p.tau_g + 0.5);
  d.force_pf = 1. / (1. + 1. / (2. * p.tau_f));
  d.acc_coef = -(1. / 3.) * p.alpha0;
  const double lam = 1. / (p.tau_f + 0.5);
  const double A = 2. * (lam - 0.5 * lam * lam);
  d.amp_j = A * p.kBT;
  d.amp_s = A * p.kBT / (1. / 3.);
  d.sqrt_amp_j = sqrt(d.amp_j);
  d.sqrt_amp_s = sqrt(d.amp_s);
  d.keys = philox_key_schedule(p.seed);
}

template <class T>
int dev_alloc(bflbm_lattice* h, T** p, size_t count) {
  CU(cudaMalloc((void**)p, count * sizeof(T)));
  h->bytes += count * sizeof(T);
  return 0;
}

int ensure_stage(bflbm_lattice* h, size_t doubles) {
  if (h->stage_doubles >= doubles) return 0;
  if (h->stage) {
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaFree(h->stage));
    h->bytes -= h->stage_doubles * sizeof(double);
    h->stage = nullptr;
    h->stage_doubles = 0;
  }
  CU(cudaMalloc((void**)&h->stage, doubles * sizeof(double)));
  h->stage_doubles = doubles;
  h->bytes += doubles * sizeof(double);
  return 0;
}

// planes per chunk so that ncomp*planes*plane doubles stay below 1 GiB
int chunk_planes(const bflbm_lattice* h, int ncomp, int extra_planes) {
  const size_t budget = (size_t)128 << 20;  // doubles (1 GiB stage)
  long long per_plane = (long long)ncomp * h->G.plane;
  long long n = (long long)(budget / (size_t)per_plane) - extra_planes;
  if (n < 1) n = 1;
  if (n > h->G.nzl) n = h->G.nzl;
  return (int)n;
}

dim3 cell_grid(const bflbm_lattice* h, int planes) { return dim3(h->grid_xy.x, h->grid_xy.y, planes); }

// ---- ghost planes of a whole-box lattice (periodic in z inside one GPU) --------------------------------
int wrap_population_ghosts(bflbm_lattice* h, double* X) {
  const Geom& G = h->G;
  CopyList L;
  L.n = 0;
  for (int s = 0; s < 2; ++s)
    for (int i = 0; i < Q; ++i) {
      const long long base = (long long)(s * Q + i) * G.comp;
      if (cz(i) == 1) {  // pulled from plane below: ghost plane 0 <- plane nzl
        L.src[L.n] = base + (long long)G.nzl * G.plane;
        L.dst[L.n] = base;
        ++L.n;
      } else if (cz(i) == -1) {  // ghost plane nzl+1 <- plane 1
        L.src[L.n] = base + G.plane;
        L.dst[L.n] = base + (long long)(G.nzl + 1) * G.plane;
        ++L.n;
      }
    }
  const int T = 256;
  dim3 grid((unsigned)((G.plane + T - 1) / T), L.n);
  k_copy_planes<<<grid, T, 0, h->stream>>>(L, X, X, G.plane);
  ++h->launches;
  CU(cudaGetLastError());
  return 0;
}
int wrap_density_ghosts(bflbm_lattice* h) {
  const Geom& G = h->G;
  CopyList L;
  L.n = 2;
  L.src[0] = 2 * (long long)G.nzl * G.plane; L.dst[0] = 0;
  L.src[1] = 2 * G.plane;                    L.dst[1] = 2 * (long long)(G.nzl + 1) * G.plane;
  const int T = 256;
  dim3 grid((unsigned)((2 * G.plane + T - 1) / T), L.n);
  k_copy_planes<<<grid, T, 0, h->stream>>>(L, (const double*)h->R, (double*)h->R, 2 * G.plane);
  ++h->launches;
  CU(cudaGetLastError());
  return 0;
}
int density_pass(bflbm_lattice* h) {
  k_density<<<cell_grid(h, h->G.nzl), h->block, 0, h->stream>>>(h->G, h->X[h->cur], h->R);
  ++h->launches;
  CU(cudaGetLastError());
  return 0;
}

// ---- halo messages ------------------------------------------------------------------------------------
// message to the neighbour on `side` (0: lower z, 1: upper z), in doubles:
//   [10][plane]   the 5+5 populations of my boundary plane that the neighbour pulls across the face
//                 (side 0: c_z = -1 of plane 0;  side 1: c_z = +1 of plane nzl-1)
//   [2*plane]     Pz: my local partial (rho,phi) sums on my boundary plane
//   [2*plane]     Ez: my contribution to (rho,phi) on the neighbour's boundary plane
// (SURVEY.md 8(e) option (ii), folded into a single message: 112 B per face cell.)
int pack_halo(bflbm_lattice* h) {
  const Geom& G = h->G;
  for (int side = 0; side < 2; ++side) {
    CopyList L;
    L.n = 0;
    const int want = side == 0 ? -1 : 1;
    const long long bplane = side == 0 ? 1 : G.nzl;  // storage index of my boundary plane
    for (int s = 0; s < 2; ++s)
      for (int i = 0; i < Q; ++i)
        if (cz(i) == want) {
          L.src[L.n] = (long long)(s * Q + i) * G.comp + bplane * G.plane;
          L.dst[L.n] = (long long)L.n * G.plane;
          ++L.n;
        }
    const int T = 256;
    dim3 grid((unsigned)((G.plane + T - 1) / T), L.n);
    k_copy_planes<<<grid, T, 0, h->stream>>>(L, h->X[h->cur], h->send[side], G.plane);
    ++h->launches;
    // density partial planes: Pz = R[boundary], Ez = R[outside ghost] (local sums written by the fold)
    CopyList D;
    D.n = 2;
    D.src[0] = 2 * bplane * G.plane;                                  D.dst[0] = 10 * G.plane;
    D.src[1] = 2 * (side == 0 ? 0 : (long long)(G.nzl + 1)) * G.plane; D.dst[1] = 12 * G.plane;
    dim3 gridd((unsigned)((2 * G.plane + T - 1) / T), D.n);
    k_copy_planes<<<gridd, T, 0, h->stream>>>(D, (const double*)h->R, h->send[side], 2 * G.plane);
    ++h->launches;
  }
  CU(cudaGetLastError());
  return 0;
}
// recv[side] holds the message the neighbour on `side` packed for me (its side 1-side message)
int unpack_halo(bflbm_lattice* h, double* const recv[2]) {
  const Geom& G = h->G;
  for (int side = 0; side < 2; ++side) {
    CopyList L;
    L.n = 0;
    // lower neighbour sent its c_z = +1 populations (its side-1 message): they go to my ghost plane 0
    const int want = side == 0 ? 1 : -1;
    const long long gplane = side == 0 ? 0 : G.nzl + 1;
    for (int s = 0; s < 2; ++s)
      for (int i = 0; i < Q; ++i)
        if (cz(i) == want) {
          L.src[L.n] = (long long)L.n * G.plane;
          L.dst[L.n] = (long long)(s * Q + i) * G.comp + gplane * G.plane;
          ++L.n;
        }
    const int T = 256;
    dim3 grid((unsigned)((G.plane + T - 1) / T), L.n);
    k_copy_planes<<<grid, T, 0, h->stream>>>(L, recv[side], h->X[h->cur], G.plane);
    ++h->launches;
    const long long bplane = side == 0 ? 1 : G.nzl;
    dim3 gridd((unsigned)((G.plane + T - 1) / T));
    k_merge_density_halo<<<gridd, T, 0, h->stream>>>(G.plane, (const double2*)(recv[side] + 10 * G.plane),
                                                     (const double2*)(recv[side] + 12 * G.plane), h->R + bplane * G.plane,
                                                     h->R + gplane * G.plane);
    ++h->launches;
  }
  CU(cudaGetLastError());
  return 0;
}

// mode 0: all planes; 1: brick-face planes + slab-face outputs; 2: the complement of 1
// 3: the four slab-face planes (first / last brick row only); 4 = 1 without 3
int fold_local(bflbm_lattice* h, int mode, cudaStream_t st = nullptr) {
  const Geom& G = h->G;
  if (!st) st = h->stream;
  const dim3 grid(h->B.bx, h->B.by, mode == 3 ? std::min(2, h->B.bz) : h->B.bz), block(h->B.tx, h->B.ty);
  if (mode == 0)      k_fold<0><<<grid, block, 0, st>>>(G, h->B, h->E[h->ecur], h->R);
  else if (mode == 1) k_fold<1><<<grid, block, 0, st>>>(G, h->B, h->E[h->ecur], h->R);
  else if (mode == 2) k_fold<2><<<grid, block, 0, st>>>(G, h->B, h->E[h->ecur], h->R);
  else if (mode == 3) k_fold<3><<<grid, block, 0, st>>>(G, h->B, h->E[h->ecur], h->R);
  else                k_fold<4><<<grid, block, 0, st>>>(G, h->B, h->E[h->ecur], h->R);
  ++h->launches;
  CU(cudaGetLastError());
  return 0;
}
// R complete on every plane (observers, diagnostics, kernels that stage every plane from R)
int ensure_full_R(bflbm_lattice* h) {
  if (!h->r_stale) return 0;
  int rc = fold_local(h, 2);
  if (rc) return rc;
  h->r_stale = false;
  return 0;
}

// rows: 0 = every brick row, 1 = the first and the last row (two z-blocks), 2 = rows 1 .. bz-2
template <bool NOISE, bool R1, bool FU>
int launch_fused_v(bflbm_lattice* h, int rows, cudaStream_t st) {
  const BrickGrid& B = h->B;
  const dim3 grid(B.bx, B.by, rows == 0 ? B.bz : (rows == 1 ? 2 : B.bz - 2)), block(B.tx, B.ty);
  const int bz0 = rows == 2 ? 1 : 0, two_ends = rows == 1;
  const double2* Ein = (h->e_valid && h->fold_in_staging) ? h->E[h->ecur] : nullptr;
  double2* Eout = h->E[1 - h->ecur];
  const PopBases XB = make_pop_bases(h->G, h->X[h->cur], h->X[1 - h->cur]);
  if (B.tx * B.ty == 128)
    k_step_fused<NOISE, R1, FU, 128><<<grid, block, fused_smem_bytes(B, R1), st>>>(h->G, B, h->dp, h->step, XB, h->R, Ein, Eout, bz0, two_ends);
  else
    k_step_fused<NOISE, R1, FU, 256><<<grid, block, fused_smem_bytes(B, R1), st>>>(h->G, B, h->dp, h->step, XB, h->R, Ein, Eout, bz0, two_ends);
  ++h->launches;
  CU(cudaGetLastError());
  return 0;
}
template <bool NOISE>
int launch_fused(bflbm_lattice* h, int rows, cudaStream_t st) {
  const bool rate1 = h->rate1_fast_path && h->dp.rate_f == 1. && h->dp.rate_g == 1.;
  const bool full = h->G.nx % h->B.tx == 0 && h->G.ny % h->B.ty == 0;
  if (rate1) return full ? launch_fused_v<NOISE, true, true>(h, rows, st) : launch_fused_v<NOISE, true, false>(h, rows, st);
  return full ? launch_fused_v<NOISE, false, true>(h, rows, st) : launch_fused_v<NOISE, false, false>(h, rows, st);
}

// collide+stream of the slab and local density partials; leaves the outgoing messages packed
int step_local(bflbm_lattice* h, bool pack = true) {
  const bool noise = h->prm.kBT > 0.;
  mark(h, 0);
  int rc0 = 0;
  // every kernel but the default one stages all planes from R; the default one only when E is not usable
  if ((h->algo != 0 || !h->e_valid || !h->fold_in_staging) && (rc0 = ensure_full_R(h))) return rc0;
  if (h->algo == 1) {
    h->e_valid = false;
    if (noise) k_step_twopass<true><<<cell_grid(h, h->G.nzl), h->block, 0, h->stream>>>(h->G, h->dp, h->step, h->X[h->cur], h->X[1 - h->cur], h->R);
    else       k_step_twopass<false><<<cell_grid(h, h->G.nzl), h->block, 0, h->stream>>>(h->G, h->dp, h->step, h->X[h->cur], h->X[1 - h->cur], h->R);
    ++h->launches;
    CU(cudaGetLastError());
    h->cur ^= 1;
    mark(h, 1);
    return 0;
  }
  // the default kernel folds the brick-interior planes itself while staging (next step): only the brick faces here
  const bool partial = h->algo == 0 && h->fold_in_staging && fold_in_staging_ok(h->G, h->B);
  int rc;
  // The overlapped schedule needs a last brick row of at least two planes: with a single plane, the slab-face fold on
  // `stream` would read the top shell of row bz-2 while `aux` is still writing it (and overwrite the R plane that row
  // stages for its top gradients).  Such slabs take the single-stream path below.
  const bool last_row_ok = h->G.nzl - (h->B.bz - 1) * h->B.lz >= 2;
  if (!h->whole_box && h->overlap && !h->profiling && partial && h->B.bz >= 3 && last_row_ok) {
    // overlapped slab step.  stream: [first+last brick row] -> [fold of the 4 slab-face planes] -> [pack] -> caller's
    // exchange -> (step_end) unpack.  aux: [interior rows] -> [fold of the other brick faces], concurrent with the exchange.
    CU(cudaEventRecord(h->ev_prev, h->stream));          // everything queued so far (previous step, uploads, observers)
    CU(cudaStreamWaitEvent(h->aux, h->ev_prev, 0));
    if ((rc = noise ? launch_fused<true>(h, 1, h->stream) : launch_fused<false>(h, 1, h->stream))) return rc;
    CU(cudaEventRecord(h->ev_ends, h->stream));
    if ((rc = noise ? launch_fused<true>(h, 2, h->aux) : launch_fused<false>(h, 2, h->aux))) return rc;
    h->cur ^= 1;
    h->ecur ^= 1;
    h->e_valid = true;
    h->r_stale = true;
    if ((rc = fold_local(h, 3))) return rc;
    if ((rc = pack_halo(h))) return rc;
    CU(cudaStreamWaitEvent(h->aux, h->ev_ends, 0));      // the brick faces next to the end rows need their extended boxes
    if ((rc = fold_local(h, 4, h->aux))) return rc;
    CU(cudaEventRecord(h->ev_interior, h->aux));
    h->interior_pending = true;
    return 0;
  }
  rc = noise ? launch_fused<true>(h, 0, h->stream) : launch_fused<false>(h, 0, h->stream);
  if (rc) return rc;
  h->cur ^= 1;
  h->ecur ^= 1;
  mark(h, 1);
  h->e_valid = partial;
  h->r_stale = partial;
  if ((rc = fold_local(h, partial ? 1 : 0))) return rc;
  mark(h, 2);
  if (pack) rc = pack_halo(h);
  mark(h, 3);
  return rc;
}

int create_common(const bflbm_params* p, int nx, int ny, int nz_global, int z0, int nzl, int device, bool whole, bflbm_lattice** out) {
  if (!out) return fail(BFLBM_ERR_ARG, "null out pointer");
  *out = nullptr;
  int rc = validate_params(p);
  if (rc) return rc;
  if (nx < 1 || ny < 1 || nz_global < 2 || nzl < 2 || z0 < 0 || z0 + nzl > nz_global)
    return fail(BFLBM_ERR_ARG, "bad lattice size nx=%d ny=%d nz=%d z0=%d nz_local=%d (need nx,ny >= 1, nz_local >= 2)", nx, ny, nz_global, z0, nzl);
  if ((double)(nzl + 2) * nx * ny * 8.0 >= 4294967296.0)
    return fail(BFLBM_ERR_ARG, "slab too large: (nz_local+2)*nx*ny*8 B must stay below 4 GiB per component (32-bit in-component offsets); "
                               "use more slabs");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(BFLBM_ERR_CUDA, "no CUDA device available: this library has no CPU path");
  if (device < 0 || device >= ndev) return fail(BFLBM_ERR_ARG, "device %d out of range (have %d)", device, ndev);
  bflbm_lattice* h = new (std::nothrow) bflbm_lattice;
  if (!h) return fail(BFLBM_ERR_ARG, "out of host memory");
  h->device = device;
  h->whole_box = whole;
  h->prm = *p;
  h->step = p->step0;
  derive(h);
  Geom& G = h->G;
  G.nx = nx; G.ny = ny; G.nzl = nzl; G.nz_global = nz_global; G.z0 = z0;
  G.plane = (long long)nx * ny;
  G.comp = (long long)(nzl + 2) * G.plane;
  int bx = 8;
  while (bx < nx && bx < 128) bx <<= 1;
  h->block = dim3(bx, 256 / bx);
  h->grid_xy = dim3(ceil_div(nx, bx), ceil_div(ny, 256 / bx));
#define TRY(x) if ((rc = (x))) { bflbm_destroy(h); return rc; }
  TRY(set_device(h));
  {
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; return fail(BFLBM_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    h->own_stream = true;
  }
  TRY(dev_alloc(h, &h->X[0], (size_t)(2 * Q) * G.comp));
  TRY(dev_alloc(h, &h->X[1], (size_t)(2 * Q) * G.comp));
  TRY(dev_alloc(h, &h->R, (size_t)G.comp));
  {
    const char* nt = getenv("BFLBM_CTA_THREADS");
    h->cta_threads = (nt && atoi(nt) == 128) ? 128 : 256;
  }
  h->B = make_brick_grid(G, 0, h->cta_threads, 32);
  // a slab wants at least 3 brick rows so that the halo exchange can overlap the interior rows (thin strong-scaling slabs)
  while (!whole && h->B.bz < 3 && h->B.lz >= 16) h->B = make_brick_grid(G, h->B.lz / 2, h->cta_threads, 32);
  TRY(dev_alloc(h, &h->E[0], brick_doubles2(h->B)));
  TRY(dev_alloc(h, &h->E[1], brick_doubles2(h->B)));
  if (!whole) {
    const char* ov = getenv("BFLBM_OVERLAP");
    h->overlap = !(ov && ov[0] == '0');
    cudaError_t e = cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_prev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_ends, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_interior, cudaEventDisableTiming);
    if (e != cudaSuccess) { bflbm_destroy(h); return fail(BFLBM_ERR_CUDA, "aux stream / events: %s", cudaGetErrorString(e)); }
  }
  {
    const char* fs = getenv("BFLBM_FOLD_IN_STAGING");
    h->fold_in_staging = !(fs && fs[0] == '0');
  }
  h->halo_doubles = (size_t)14 * G.plane;
  for (int s = 0; s < 2; ++s) {
    TRY(dev_alloc(h, &h->send[s], h->halo_doubles));
    if (!whole) TRY(dev_alloc(h, &h->recv[s], h->halo_doubles));
  }
  h->diag_blocks = (size_t)h->grid_xy.x * h->grid_xy.y * nzl;
  TRY(dev_alloc(h, &h->diag_partial, h->diag_blocks * NDIAG));
  TRY(dev_alloc(h, &h->diag_count, (size_t)1));
  {
    const char* r1 = getenv("BFLBM_RATE1");
    h->rate1_fast_path = !(r1 && r1[0] == '0');
    cudaError_t e = set_fused_smem();
    if (e != cudaSuccess) { bflbm_destroy(h); return fail(BFLBM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); }
  }
#undef TRY
  *out = h;
  return 0;
}

// after X planes 1..nzl are set (whole box): make ghosts, densities consistent
int finish_init(bflbm_lattice* h) {
  h->step = h->prm.step0;
  h->initialized = true;
  h->e_valid = false;  // R was (or is about to be) rebuilt from the populations; E is void
  h->r_stale = false;
  return 0;
}

int run_init(bflbm_lattice* h, const InitSpec& S) {
  int rc = set_device(h);
  if (rc) return rc;
  k_init<<<cell_grid(h, h->G.nzl + 2), h->block, 0, h->stream>>>(h->G, S, h->X[h->cur], h->R);
  ++h->launches;
  CU(cudaGetLastError());
  return finish_init(h);
}

// generic chunked observer -> host (or device) array of ncomp components
template <int MODE>
int observe(bflbm_lattice* h, int ncomp, double* out, bool out_is_device, bool cell_major) {
  CHECK_H(h);
  if (!out) return fail(BFLBM_ERR_ARG, "null output buffer");
  if (!h->initialized) return fail(BFLBM_ERR_STATE, "lattice not initialised");
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = ensure_full_R(h))) return rc;
  const Geom& G = h->G;
  const bool noise = h->prm.kBT > 0.;
  const int cp = chunk_planes(h, ncomp, 0);
  if ((rc = ensure_stage(h, (size_t)ncomp * cp * G.plane))) return rc;
  for (int zlo = 0; zlo < G.nzl; zlo += cp) {
    const int zc = std::min(cp, G.nzl - zlo);
    if (noise) k_observe<MODE, true><<<cell_grid(h, zc), h->block, 0, h->stream>>>(G, h->dp, h->step, zlo, h->X[h->cur], h->R, h->stage);
    else       k_observe<MODE, false><<<cell_grid(h, zc), h->block, 0, h->stream>>>(G, h->dp, h->step, zlo, h->X[h->cur], h->R, h->stage);
    ++h->launches;
    CU(cudaGetLastError());
    const cudaMemcpyKind kind = out_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (cell_major) {
      CU(cudaMemcpyAsync(out + (size_t)zlo * G.plane * ncomp, h->stage, (size_t)zc * G.plane * ncomp * sizeof(double), kind, h->stream));
    } else {
      CU(cudaMemcpy2DAsync(out + (size_t)zlo * G.plane, (size_t)G.nzl * G.plane * sizeof(double), h->stage,
                           (size_t)zc * G.plane * sizeof(double), (size_t)zc * G.plane * sizeof(double), ncomp, kind, h->stream));
    }
    // the stage is reused by the next chunk: stream order (kernel after copy) protects it, no host sync per chunk
  }
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int upload_populations(bflbm_lattice* h, const double* f, const double* g, bool ghosted) {
  const Geom& G = h->G;
  int rc;
  // host planes are the sources: [0, nzl) of a whole box, [-1, nzl] of a slab (ghosted arrays, host plane index + 1)
  const int zfirst = ghosted ? -1 : 0, nsrc = ghosted ? G.nzl + 2 : G.nzl;
  const size_t budget = (size_t)128 << 20;  // doubles (1 GiB): a few large copies per component instead of one per plane
  int cp = (int)std::max<size_t>(1, std::min<size_t>((size_t)nsrc, budget / ((size_t)(2 * Q) * (size_t)G.plane)));
  if ((rc = ensure_stage(h, (size_t)(2 * Q) * cp * G.plane))) return rc;
  const double* src[2] = {f, g};
  const size_t spitch = (size_t)nsrc * G.plane * sizeof(double);
  for (int z = 0; z < nsrc; z += cp) {
    const int zc = std::min(cp, nsrc - z);
    const size_t dpitch = (size_t)zc * G.plane * sizeof(double);
    for (int s = 0; s < 2; ++s)
      CU(cudaMemcpy2DAsync(h->stage + (size_t)s * Q * zc * G.plane, dpitch, src[s] + (size_t)z * G.plane, spitch, dpitch, Q,
                           cudaMemcpyHostToDevice, h->stream));
    // stream order protects the stage: the next chunk's copies start after this kernel has read it (no host sync)
    k_scatter_populations<<<cell_grid(h, zc), h->block, 0, h->stream>>>(G, zfirst + z, ghosted ? 0 : 1, h->stage, h->X[h->cur]);
    ++h->launches;
    CU(cudaGetLastError());
  }
  CU(cudaStreamSynchronize(h->stream));  // the caller's host buffers are free again
  return 0;
}

}  // namespace

extern "C" {

int bflbm_params_default(bflbm_params* p) {
  if (!p) return fail(BFLBM_ERR_ARG, "null params");
  p->kBT = 0.; p->tau_f = 0.5; p->tau_g = 0.5; p->alpha0 = 4.; p->alpha1 = 0.; p->kappa = 4.;
  p->rho_lo = 0.; p->rho_hi = 1.; p->seed = 12345ull; p->step0 = 0;
  return 0;
}

int bflbm_create(const bflbm_params* p, int nx, int ny, int nz, int device, bflbm_lattice** out) {
  return create_common(p, nx, ny, nz, 0, nz, device, true, out);
}
int bflbm_create_slab(const bflbm_params* p, int nx, int ny, int nz_global, int z0, int nz_local, int device, bflbm_lattice** out) {
  return create_common(p, nx, ny, nz_global, z0, nz_local, device, false, out);
}

int bflbm_destroy(bflbm_lattice* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->aux) { cudaStreamSynchronize(h->aux); cudaStreamDestroy(h->aux); }
  if (h->ev_prev) cudaEventDestroy(h->ev_prev);
  if (h->ev_ends) cudaEventDestroy(h->ev_ends);
  if (h->ev_interior) cudaEventDestroy(h->ev_interior);
  cudaFree(h->X[0]); cudaFree(h->X[1]); cudaFree(h->R); cudaFree(h->E[0]); cudaFree(h->E[1]);
  for (int s = 0; s < 2; ++s) { cudaFree(h->send[s]); cudaFree(h->recv[s]); }
  cudaFree(h->stage); cudaFree(h->diag_partial); cudaFree(h->diag_count);
  for (int i = 0; i < 5; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int bflbm_set_params(bflbm_lattice* h, const bflbm_params* p) {
  CHECK_H(h);
  int rc = validate_params(p);
  if (rc) return rc;
  const long long keep = h->step;
  h->prm = *p;
  derive(h);
  h->step = h->initialized ? keep : p->step0;
  return 0;
}
int bflbm_get_params(const bflbm_lattice* h, bflbm_params* p) {
  CHECK_H(h);
  if (!p) return fail(BFLBM_ERR_ARG, "null params");
  *p = h->prm;
  return 0;
}
int bflbm_set_stream(bflbm_lattice* h, void* s) {
  CHECK_H(h);
  int rc = set_device(h);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->stream));
  if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
  if (s) h->stream = (cudaStream_t)s;
  else {
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  return 0;
}
int bflbm_set_algorithm(bflbm_lattice* h, int algo) {
  CHECK_H(h);
  if (algo < 0 || algo > 1) return fail(BFLBM_ERR_ARG, "algorithm must be 0 (fused one-pass) or 1 (two-pass)");
  if (algo == 1 && !h->whole_box) return fail(BFLBM_ERR_ARG, "the two-pass algorithm supports whole-box lattices only");
  h->algo = algo;
  return bflbm_set_tiling(h, h->lz_request);  // the brick shape depends on the kernel
}

int bflbm_set_tiling(bflbm_lattice* h, int brick_lz) {
  CHECK_H(h);
  if (brick_lz < 0) return fail(BFLBM_ERR_ARG, "brick height must be >= 0");
  h->lz_request = brick_lz;
  int rc = set_device(h);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->stream));
  if (h->initialized && (rc = ensure_full_R(h))) return rc;  // E is about to change shape: R takes over
  CU(cudaStreamSynchronize(h->stream));
  h->e_valid = false;
  const BrickGrid nb = make_brick_grid(h->G, brick_lz, h->cta_threads, 32);
  if (brick_doubles2(nb) != brick_doubles2(h->B)) {
    for (int k = 0; k < 2; ++k) {
      CU(cudaFree(h->E[k]));
      h->bytes -= brick_doubles2(h->B) * sizeof(double2);
      h->E[k] = nullptr;
      if ((rc = dev_alloc(h, &h->E[k], brick_doubles2(nb)))) return rc;
    }
  }
  h->B = nb;
  return 0;
}

int bflbm_init_mixture(bflbm_lattice* h) {
  CHECK_H(h);
  InitSpec S{0, h->prm.rho_lo, h->prm.rho_hi, h->prm.kappa, 0., 0.};
  return run_init(h, S);
}
int bflbm_init_stripe(bflbm_lattice* h, double frac) {
  CHECK_H(h);
  InitSpec S{1, h->prm.rho_lo, h->prm.rho_hi, h->prm.kappa, frac, 0.};
  return run_init(h, S);
}
int bflbm_init_droplet(bflbm_lattice* h, double radius) {
  CHECK_H(h);
  InitSpec S{2, h->prm.rho_lo, h->prm.rho_hi, h->prm.kappa, 0., radius};
  return run_init(h, S);
}
int bflbm_init_from_populations(bflbm_lattice* h, const double* f, const double* g) {
  CHECK_H(h);
  if (!f || !g) return fail(BFLBM_ERR_ARG, "null population buffer");
  if (!h->whole_box) return fail(BFLBM_ERR_ARG, "slab lattices need bflbm_init_from_populations_slab");
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = upload_populations(h, f, g, false))) return rc;
  // same partial-sum + self-exchange path a slab takes, so the result does not depend on the slab count
  if ((rc = finish_init(h))) return rc;
  if ((rc = bflbm_halo_refresh_begin(h))) return rc;
  return bflbm_halo_refresh_end(h);
}
int bflbm_init_from_populations_slab(bflbm_lattice* h, const double* f, const double* g) {
  CHECK_H(h);
  if (!f || !g) return fail(BFLBM_ERR_ARG, "null population buffer");
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = upload_populations(h, f, g, true))) return rc;
  if ((rc = finish_init(h))) return rc;
  // the ghost planes of X and the densities are completed by bflbm_halo_refresh_begin / exchange / _end,
  // which the caller must run next (done here for a whole box, whose neighbour is itself)
  if (h->whole_box) {
    if ((rc = bflbm_halo_refresh_begin(h))) return rc;
    return bflbm_halo_refresh_end(h);
  }
  return 0;
}

int bflbm_step(bflbm_lattice* h, int nsteps) {
  CHECK_H(h);
  if (!h->initialized) return fail(BFLBM_ERR_STATE, "bflbm_step before init");
  if (!h->whole_box) return fail(BFLBM_ERR_STATE, "slab lattice: use bflbm_step_begin / exchange / bflbm_step_end");
  if (nsteps < 0) return fail(BFLBM_ERR_ARG, "nsteps < 0");
  int rc = set_device(h);
  if (rc) return rc;
  for (int s = 0; s < nsteps; ++s) {
    if ((rc = step_local(h, /*pack=*/false))) return rc;
    if (h->algo == 1) {
      if ((rc = wrap_population_ghosts(h, h->X[h->cur]))) return rc;
      mark(h, 2);  // two-pass: interval 2 = population ghost wrap, interval 1 -> density pass below
      if ((rc = density_pass(h))) return rc;
      mark(h, 3);
      if ((rc = wrap_density_ghosts(h))) return rc;
    } else {
      // periodic self-exchange in one launch (what pack_halo + unpack_halo of my own messages would do)
      k_wrap_whole_box<<<(unsigned)((h->G.plane + 255) / 256), 256, 0, h->stream>>>(h->G, h->X[h->cur], h->R);
      ++h->launches;
      CU(cudaGetLastError());
    }
    mark(h, 4);
    profile_collect(h);
    ++h->step;
  }
  return 0;
}
int bflbm_step_begin(bflbm_lattice* h) {
  CHECK_H(h);
  if (!h->initialized) return fail(BFLBM_ERR_STATE, "bflbm_step_begin before init");
  if (h->algo != 0) return fail(BFLBM_ERR_STATE, "slab stepping needs the fused algorithm");
  int rc = set_device(h);
  if (rc) return rc;
  return step_local(h);
}
int bflbm_step_end(bflbm_lattice* h) {
  CHECK_H(h);
  int rc = set_device(h);
  if (rc) return rc;
  if (h->interior_pending) {  // join the interior rows before anything else touches the lattice
    CU(cudaStreamWaitEvent(h->stream, h->ev_interior, 0));
    h->interior_pending = false;
  }
  double* const self[2] = {h->send[1], h->send[0]};
  if ((rc = unpack_halo(h, h->whole_box ? self : h->recv))) return rc;
  mark(h, 4);
  profile_collect(h);  // interval 3 (pack end -> unpack end) contains the caller's exchange
  ++h->step;
  return 0;
}
int bflbm_halo_refresh_begin(bflbm_lattice* h) {
  CHECK_H(h);
  if (!h->initialized) return fail(BFLBM_ERR_STATE, "lattice not initialised");
  int rc = set_device(h);
  if (rc) return rc;
  h->e_valid = false;  // R is rebuilt from the populations below
  h->r_stale = false;
  // local density partials straight from the populations: P = sum over owned source planes only
  k_density_partial<<<cell_grid(h, h->G.nzl + 2), h->block, 0, h->stream>>>(h->G, h->X[h->cur], h->R);
  ++h->launches;
  CU(cudaGetLastError());
  return pack_halo(h);
}
int bflbm_halo_refresh_end(bflbm_lattice* h) {
  CHECK_H(h);
  int rc = set_device(h);
  if (rc) return rc;
  double* const self[2] = {h->send[1], h->send[0]};
  return unpack_halo(h, h->whole_box ? self : h->recv);
}
size_t bflbm_halo_doubles(const bflbm_lattice* h) { return h ? h->halo_doubles : 0; }
void* bflbm_halo_send_buffer(bflbm_lattice* h, int side) { return (h && (side == 0 || side == 1)) ? h->send[side] : nullptr; }
void* bflbm_halo_recv_buffer(bflbm_lattice* h, int side) { return (h && (side == 0 || side == 1)) ? h->recv[side] : nullptr; }

int bflbm_sync(bflbm_lattice* h) {
  CHECK_H(h);
  int rc = set_device(h);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}
long long bflbm_step_count(const bflbm_lattice* h) { return h ? h->step : -1; }
int bflbm_get_dims(const bflbm_lattice* h, int* nx, int* ny, int* nz_local, int* z0, int* nz_global) {
  CHECK_H(h);
  if (nx) *nx = h->G.nx;
  if (ny) *ny = h->G.ny;
  if (nz_local) *nz_local = h->G.nzl;
  if (z0) *z0 = h->G.z0;
  if (nz_global) *nz_global = h->G.nz_global;
  return 0;
}

int bflbm_get_populations(bflbm_lattice* h, double* f, double* g) {
  CHECK_H(h);
  if (!f || !g) return fail(BFLBM_ERR_ARG, "null output buffer");
  // one observer pass produces both species; split into the two host arrays
  if (!h->initialized) return fail(BFLBM_ERR_STATE, "lattice not initialised");
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = ensure_full_R(h))) return rc;
  const Geom& G = h->G;
  const int cp = chunk_planes(h, 2 * Q, 0);
  if ((rc = ensure_stage(h, (size_t)(2 * Q) * cp * G.plane))) return rc;
  for (int zlo = 0; zlo < G.nzl; zlo += cp) {
    const int zc = std::min(cp, G.nzl - zlo);
    k_observe<OBS_POP, false><<<cell_grid(h, zc), h->block, 0, h->stream>>>(G, h->dp, h->step, zlo, h->X[h->cur], h->R, h->stage);
    ++h->launches;
    CU(cudaGetLastError());
    double* outs[2] = {f, g};
    for (int s = 0; s < 2; ++s)
      CU(cudaMemcpy2DAsync(outs[s] + (size_t)zlo * G.plane, (size_t)G.nzl * G.plane * sizeof(double), h->stage + (size_t)s * Q * zc * G.plane,
                           (size_t)zc * G.plane * sizeof(double), (size_t)zc * G.plane * sizeof(double), Q, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  return 0;
}
int bflbm_get_populations_device(bflbm_lattice* h, double* dev_f, double* dev_g) {
  CHECK_H(h);
  if (!dev_f || !dev_g) return fail(BFLBM_ERR_ARG, "null output buffer");
  if (!h->initialized) return fail(BFLBM_ERR_STATE, "lattice not initialised");
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = ensure_full_R(h))) return rc;
  const Geom& G = h->G;
  const int cp = chunk_planes(h, 2 * Q, 0);
  if ((rc = ensure_stage(h, (size_t)(2 * Q) * cp * G.plane))) return rc;
  for (int zlo = 0; zlo < G.nzl; zlo += cp) {
    const int zc = std::min(cp, G.nzl - zlo);
    k_observe<OBS_POP, false><<<cell_grid(h, zc), h->block, 0, h->stream>>>(G, h->dp, h->step, zlo, h->X[h->cur], h->R, h->stage);
    ++h->launches;
    CU(cudaGetLastError());
    double* outs[2] = {dev_f, dev_g};
    for (int s = 0; s < 2; ++s)
      CU(cudaMemcpy2DAsync(outs[s] + (size_t)zlo * G.plane, (size_t)G.nzl * G.plane * sizeof(double), h->stage + (size_t)s * Q * zc * G.plane,
                           (size_t)zc * G.plane * sizeof(double), (size_t)zc * G.plane * sizeof(double), Q, cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  return 0;
}
int bflbm_get_hydrovars(bflbm_lattice* h, double* out22) { return observe<OBS_HYDRO>(h, BFLBM_NHYDRO, out22, false, false); }
int bflbm_get_hydrovars_bar(bflbm_lattice* h, double* out9) { return observe<OBS_HBAR>(h, BFLBM_NHYDRO_BAR, out9, false, false); }
int bflbm_get_hydrovars_device(bflbm_lattice* h, double* o) { return observe<OBS_HYDRO>(h, BFLBM_NHYDRO, o, true, false); }
int bflbm_get_hydrovars_bar_device(bflbm_lattice* h, double* o) { return observe<OBS_HBAR>(h, BFLBM_NHYDRO_BAR, o, true, false); }
int bflbm_get_noise(bflbm_lattice* h, double* fn, double* gn) {
  CHECK_H(h);
  if (!fn || !gn) return fail(BFLBM_ERR_ARG, "null output buffer");
  std::vector<double> tmp;
  const size_t n = (size_t)h->G.nzl * h->G.plane;
  try { tmp.resize(2 * Q * n); } catch (...) { return fail(BFLBM_ERR_ARG, "out of host memory"); }
  int rc = observe<OBS_NOISE>(h, 2 * Q, tmp.data(), false, false);
  if (rc) return rc;
  memcpy(fn, tmp.data(), Q * n * sizeof(double));
  memcpy(gn, tmp.data() + Q * n, Q * n * sizeof(double));
  return 0;
}
int bflbm_get_normals(bflbm_lattice* h, double* out33) { return observe<OBS_NORMALS>(h, BFLBM_NNORMALS, out33, false, true); }

static int run_diag(bflbm_lattice* h, double sums[NDIAG], unsigned long long* bad) {
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = ensure_full_R(h))) return rc;
  CU(cudaMemsetAsync(h->diag_count, 0, sizeof(unsigned long long), h->stream));
  k_diag<<<cell_grid(h, h->G.nzl), h->block, 0, h->stream>>>(h->G, h->R, h->diag_partial, h->diag_count);
  ++h->launches;
  CU(cudaGetLastError());
  std::vector<double> part(h->diag_blocks * NDIAG);
  CU(cudaMemcpyAsync(part.data(), h->diag_partial, part.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(bad, h->diag_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  for (int k = 0; k < NDIAG; ++k) sums[k] = 0.;
  for (size_t b = 0; b < h->diag_blocks; ++b)
    for (int k = 0; k < NDIAG; ++k) sums[k] += part[b * NDIAG + k];
  return 0;
}
int bflbm_center_of_mass(bflbm_lattice* h, double* com3, double* sums4) {
  CHECK_H(h);
  if (!h->initialized) return fail(BFLBM_ERR_STATE, "lattice not initialised");
  double s[NDIAG];
  unsigned long long bad;
  int rc = run_diag(h, s, &bad);
  if (rc) return rc;
  if (com3) { com3[0] = s[2] / s[0]; com3[1] = s[3] / s[0]; com3[2] = s[4] / s[0]; }
  if (sums4) { sums4[0] = s[0]; sums4[1] = s[2]; sums4[2] = s[3]; sums4[3] = s[4]; }
  return 0;
}
int bflbm_total_mass(bflbm_lattice* h, double* mr, double* mp) {
  CHECK_H(h);
  if (!h->initialized) return fail(BFLBM_ERR_STATE, "lattice not initialised");
  double s[NDIAG];
  unsigned long long bad;
  int rc = run_diag(h, s, &bad);
  if (rc) return rc;
  if (mr) *mr = s[0];
  if (mp) *mp = s[1];
  return 0;
}
int bflbm_second_moments(bflbm_lattice* h, double* sums10) {
  CHECK_H(h);
  if (!sums10) return fail(BFLBM_ERR_ARG, "null output");
  if (!h->initialized) return fail(BFLBM_ERR_STATE, "lattice not initialised");
  double s[NDIAG];
  unsigned long long bad;
  int rc = run_diag(h, s, &bad);
  if (rc) return rc;
  sums10[0] = s[0];
  for (int k = 1; k < 10; ++k) sums10[k] = s[k + 1];
  return 0;
}
// eigenvalues of a symmetric 3x3 matrix {xx, yy, zz, xy, xz, yz}, ascending (trigonometric closed form)
static void sym3_eigenvalues(const double c[6], double e[3]) {
  const double p1 = c[3] * c[3] + c[4] * c[4] + c[5] * c[5];
  const double q = (c[0] + c[1] + c[2]) / 3.;
  const double p2 = (c[0] - q) * (c[0] - q) + (c[1] - q) * (c[1] - q) + (c[2] - q) * (c[2] - q) + 2. * p1;
  if (p2 <= 0.) { e[0] = e[1] = e[2] = q; return; }
  const double p = sqrt(p2 / 6.);
  const double b[6] = {(c[0] - q) / p, (c[1] - q) / p, (c[2] - q) / p, c[3] / p, c[4] / p, c[5] / p};
  const double det = b[0] * (b[1] * b[2] - b[5] * b[5]) - b[3] * (b[3] * b[2] - b[5] * b[4]) + b[4] * (b[3] * b[5] - b[1] * b[4]);
  const double r = std::max(-1., std::min(1., det / 2.));
  const double phi = acos(r) / 3.;
  e[2] = q + 2. * p * cos(phi);
  e[0] = q + 2. * p * cos(phi + 2. * M_PI / 3.);
  e[1] = 3. * q - e[0] - e[2];
}
int bflbm_droplet_covariance(bflbm_lattice* h, double* com3, double* cov6, double* eig3) {
  CHECK_H(h);
  if (!h->whole_box) return fail(BFLBM_ERR_ARG, "slab lattice: combine bflbm_second_moments of all slabs instead");
  double m[10];
  int rc = bflbm_second_moments(h, m);
  if (rc) return rc;
  const double M = m[0], cx = m[1] / M, cy = m[2] / M, cz = m[3] / M;
  const double c[6] = {m[4] / M - cx * cx, m[5] / M - cy * cy, m[6] / M - cz * cz, m[7] / M - cx * cy, m[8] / M - cx * cz, m[9] / M - cy * cz};
  if (com3) { com3[0] = cx; com3[1] = cy; com3[2] = cz; }
  if (cov6) for (int k = 0; k < 6; ++k) cov6[k] = c[k];
  if (eig3) sym3_eigenvalues(c, eig3);
  return 0;
}
int bflbm_check_nan(bflbm_lattice* h, long long* count) {
  CHECK_H(h);
  if (!h->initialized) return fail(BFLBM_ERR_STATE, "lattice not initialised");
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = ensure_full_R(h))) return rc;
  const Geom& G = h->G;
  const bool noise = h->prm.kBT > 0.;
  const int cp = chunk_planes(h, BFLBM_NHYDRO, 0);
  if ((rc = ensure_stage(h, (size_t)BFLBM_NHYDRO * cp * G.plane))) return rc;
  CU(cudaMemsetAsync(h->diag_count, 0, sizeof(unsigned long long), h->stream));
  for (int zlo = 0; zlo < G.nzl; zlo += cp) {
    const int zc = std::min(cp, G.nzl - zlo);
    if (noise) k_observe<OBS_HYDRO, true><<<cell_grid(h, zc), h->block, 0, h->stream>>>(G, h->dp, h->step, zlo, h->X[h->cur], h->R, h->stage);
    else       k_observe<OBS_HYDRO, false><<<cell_grid(h, zc), h->block, 0, h->stream>>>(G, h->dp, h->step, zlo, h->X[h->cur], h->R, h->stage);
    const long long n = (long long)BFLBM_NHYDRO * zc * G.plane;
    k_count_nonfinite<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->stage, n, h->diag_count);
    h->launches += 2;
    CU(cudaGetLastError());
  }
  unsigned long long bad = 0;
  CU(cudaMemcpyAsync(&bad, h->diag_count, sizeof bad, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (count) *count = (long long)bad;
  if (bad) return fail(BFLBM_ERR_NAN, "%llu non-finite values in the hydrodynamic fields at step %lld", bad, h->step);
  return 0;
}

int bflbm_set_profiling(bflbm_lattice* h, int on) {
  CHECK_H(h);
  int rc = set_device(h);
  if (rc) return rc;
  if (on && !h->ev[0])
    for (int i = 0; i < 5; ++i) CU(cudaEventCreate(&h->ev[i]));
  h->profiling = on != 0;
  for (int i = 0; i < 4; ++i) h->prof_ms[i] = 0.;
  h->prof_steps = 0;
  return 0;
}
int bflbm_get_profile(bflbm_lattice* h, double ms4[4], long long* steps) {
  CHECK_H(h);
  if (!ms4) return fail(BFLBM_ERR_ARG, "null output");
  for (int i = 0; i < 4; ++i) ms4[i] = h->prof_ms[i];
  if (steps) *steps = h->prof_steps;
  return 0;
}

long long bflbm_kernel_launches(const bflbm_lattice* h) { return h ? h->launches : 0; }
size_t bflbm_device_bytes(const bflbm_lattice* h) { return h ? h->bytes : 0; }

int bflbm_debug_philox(const unsigned int ctr[4], const unsigned int key[2], unsigned int out[4]) {
  if (!ctr || !key || !out) return fail(BFLBM_ERR_ARG, "null argument");
  uint4* d = nullptr;
  CU(cudaMalloc((void**)&d, sizeof(uint4)));
  k_philox_test<<<1, 1>>>(make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]), make_uint2(key[0], key[1]), d);
  uint4 r;
  cudaError_t e = cudaMemcpy(&r, d, sizeof r, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(BFLBM_ERR_CUDA, "philox test: %s", cudaGetErrorString(e));
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
  return 0;
}

const char* bflbm_last_error(void) { return g_err.c_str(); }
const char* bflbm_version(void) { return "bflbm-b200 0.1 (sm_100a, fp64, D3Q19 binary fluctuating)"; }

}  // extern "C"
