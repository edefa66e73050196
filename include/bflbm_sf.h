/* bflbm_sf.h -- on-GPU structure-factor accumulator for the fluctuating binary D3Q19 lattice.
 *
 * Replaces FHDeX's StructFact as the reference driver uses it:
 *   StructFact structFact(ba, dm, var_names, var_scaling, pairA, pairB);        main_run_job.cpp:299-310
 *   structFact.FortStructure(hydrovs, 0);      every out_SF_step steps          main_run_job.cpp:342-349
 *   structFact.WritePlotFile(step, time, root, zero_avg = 1);                   main_run_job.cpp:50-54
 * with the transform convention the reference states in-tree (FHDeX itself is third-party and not in the tree):
 *   r2c DFT of every variable, 1/sqrt(N) per transform, Hermitian completion, fftshift, k = 0 bin zeroed
 *   (AMReX_DFT.H:19-132, 138-183).
 * The 22 hydrodynamic fields never leave the GPU: observer kernel -> cuFFT D2Z -> running sums of A_k conj(B_k).
 *
 * A separate shared library (libbflbm_sf.so, links cuFFT) on top of the public C ABI of libbflbm.so, so that the
 * step library itself carries no FFT dependency.  The box may be one whole-box lattice or the slabs of a bflbm_multi (one
 * process, several GPUs: main_run_job.cpp:342-349 runs FortStructure on the distributed MultiFab): every slab then copies its
 * planes of the 22 fields peer-to-peer into one assembled array on the first slab's GPU, where the transforms and sums live.
 */
#ifndef BFLBM_SF_H_
#define BFLBM_SF_H_

#include "bflbm.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bflbm_sf bflbm_sf; /* opaque */

/* pairA/pairB: npairs component indices into hydrovs (0..21, VariableNames order), like StructFact's pair lists.
 * var_scaling: npairs factors applied to the accumulated pair (NULL = all 1, as at main_run_job.cpp:306-308). */
int bflbm_sf_create(bflbm_lattice* h, int npairs, const int* pairA, const int* pairB, const double* var_scaling, bflbm_sf** out);
/* the same for a box decomposed over the GPUs of this process */
int bflbm_sf_create_multi(bflbm_multi* m, int npairs, const int* pairA, const int* pairB, const double* var_scaling, bflbm_sf** out);
int bflbm_sf_destroy(bflbm_sf* s);
/* StructFact::FortStructure(hydrovs, reset = 0): transform the current hydrovs and add A_k conj(B_k) of every pair */
int bflbm_sf_accumulate(bflbm_sf* s);
int bflbm_sf_reset(bflbm_sf* s);
long long bflbm_sf_samples(const bflbm_sf* s);
/* What StructFact::WritePlotFile writes: the sample mean of Re / Im of every pair on the full (shifted) k grid,
 * host arrays of shape (npairs, nz, ny, nx), k = 0 at (nz/2, ny/2, nx/2); zero_avg != 0 zeroes that bin.
 * imag may be NULL. */
int bflbm_sf_get(bflbm_sf* s, int zero_avg, double* real, double* imag);

#ifdef __cplusplus
}
#endif
#endif /* BFLBM_SF_H_ */
