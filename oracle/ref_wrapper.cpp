// TEST INFRASTRUCTURE ONLY -- the checker, never the product.
//
// Builds the reference's own hot-path headers, UNMODIFIED and read from
// /root/reference at compile time (-I/root/reference), over the test-only
// AMReX stand-in in oracle/shim/amrex_shim.H, and exposes them through a tiny
// C ABI for ctypes (oracle/oracle.py).  Output: oracle/_ref/libbflbm_ref.so
// (git-ignored; built by oracle/Makefile).  No reference source is copied.
//
// The three sibling headers that LBM_binary.H includes (LBM_binary.H:8-10)
// pull in AMReX/FHDeX/Eigen; they are switched off through their own include
// guards (AMReX_FileIO.H:1-2, Debug.H:1-2, LBM_hydrovs.H:1-2).  The only
// symbol the hot path needs from them is update_com (LBM_hydrovs.H:26-60),
// restated below (its result is dead for the dynamics: USE_REF_STATE is
// commented out, LBM_binary.H:12).
#include "shim/amrex_shim.H"

#define LBM_IO_
#define LBM_DEB_
#define LBM_HYDRO_

// restatement of LBM_hydrovs.H:26-60 (centre of mass of component 0)
inline void update_com(const Geometry& geom, RealVect& pos_com, MultiFab& hydrovsbar, bool /*printDetails*/ = false) {
  (void)geom;
  Array4<Real> a = hydrovsbar.array4();
  const int nx = hydrovsbar.vbox.length(0), ny = hydrovsbar.vbox.length(1), nz = hydrovsbar.vbox.length(2);
  Real mass = 0., sx = 0., sy = 0., sz = 0.;
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        const Real r = a(i, j, k, 0);
        mass += r; sx += r * i; sy += r * j; sz += r * k;
      }
  pos_com[0] = sx / mass; pos_com[1] = sy / mass; pos_com[2] = sz / mass;
}

#include "LBM_binary.H"  // resolved from /root/reference by -I; includes LBM_d3q19.H

namespace {
struct RefLattice {
  Geometry geom;
  BoxArray ba;
  DistributionMapping dm;
  MultiFab fold, fnew, gold, gnew, hydrovs, hydrovsbar, fnoise, gnoise, rho_eq, phi_eq, rhot_eq;
  Vector<RealVect> com_ref;
  int nx, ny, nz;
};

void copy_out(const MultiFab& mf, int c0, int nc, double* out) {
  Array4<Real> a = mf.array4();
  const int nx = mf.vbox.length(0), ny = mf.vbox.length(1), nz = mf.vbox.length(2);
  size_t o = 0;
  for (int n = 0; n < nc; ++n)
    for (int k = 0; k < nz; ++k)
      for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) out[o++] = a(i, j, k, c0 + n);
}
void copy_in(MultiFab& mf, int nc, const double* in) {
  Array4<Real> a = mf.array4();
  const int nx = mf.vbox.length(0), ny = mf.vbox.length(1), nz = mf.vbox.length(2);
  size_t o = 0;
  for (int n = 0; n < nc; ++n)
    for (int k = 0; k < nz; ++k)
      for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) a(i, j, k, n) = in[o++];
}
}  // namespace

extern "C" {

// nghost = 2, 19/19/22/15/19/19 components: main_run_job.cpp:145-147, 205-212
void* ref_create(int nx, int ny, int nz) {
  RefLattice* L = new RefLattice;
  L->nx = nx; L->ny = ny; L->nz = nz;
  Box domain(IntVect(0, 0, 0), IntVect(nx - 1, ny - 1, nz - 1));
  L->geom = Geometry(domain);
  L->ba.domain = domain;
  const int ng = 2;
  L->fold.define(domain, nvel, ng);  L->fnew.define(domain, nvel, ng);
  L->gold.define(domain, nvel, ng);  L->gnew.define(domain, nvel, ng);
  L->hydrovs.define(domain, 22, ng); L->hydrovsbar.define(domain, 15, ng);
  L->fnoise.define(domain, nvel, ng); L->gnoise.define(domain, nvel, ng);
  L->rho_eq.define(domain, 1, ng);   L->phi_eq.define(domain, 1, ng);
  L->rhot_eq.define(domain, 1, ng);  L->rhot_eq.setVal(1.);
  for (int i = 0; i < 3; ++i) L->com_ref.push_back(RealVect(nx / 2., nx / 2., nx / 2.));  // main_run_job.cpp:117-119
  return L;
}
void ref_destroy(void* h) { delete static_cast<RefLattice*>(h); }

// run-time globals of the reference: LBM_d3q19.H:10, LBM_binary.H:17-30
void ref_set_params(double kBT_, double tau_f_, double tau_g_, double alpha0_, double alpha1_, double kappa_) {
  ::kBT = kBT_; ::tau_f = tau_f_; ::tau_g = tau_g_; ::alpha0 = alpha0_; ::alpha1 = alpha1_; ::kappa = kappa_;
}
// compile-time constants of the reference (LBM_binary.H:23-26), for the harness to read back
double ref_rho_lo() { return ::rho_lo; }
double ref_rho_hi() { return ::rho_hi; }

// rng: mode 0 = serial mt19937(seed); mode 1 = injected normals (33 per cell,
// cell-major, reference draw order); mode 2 = per-thread streams (timing).
void ref_set_rng(int mode, unsigned long seed, const double* injected) {
  ShimRng& r = shim_rng();
  r.mode = mode; r.injected = injected; r.per_cell = 33;
  amrex::InitRandom(seed);
}

void ref_init_mixture(void* h) {
  RefLattice* L = static_cast<RefLattice*>(h);
  LBM_init_mixture(L->geom, L->fold, L->gold, L->hydrovs, L->hydrovsbar, L->fnoise, L->gnoise, L->rho_eq, L->phi_eq, L->rhot_eq);
}
void ref_init_stripe(void* h, double frac) {
  RefLattice* L = static_cast<RefLattice*>(h);
  LBM_init_stripe(frac, L->geom, L->fold, L->gold, L->hydrovs, L->hydrovsbar, L->fnoise, L->gnoise, L->rho_eq, L->phi_eq, L->rhot_eq);
}
void ref_init_droplet(void* h, double radius) {
  RefLattice* L = static_cast<RefLattice*>(h);
  LBM_init_droplet(radius, L->geom, L->fold, L->gold, L->hydrovs, L->hydrovsbar, L->fnoise, L->gnoise, L->rho_eq, L->phi_eq, L->rhot_eq);
}
// restart entry LBM_init (LBM_binary.H:632-661); f0,g0: 19 comps, valid cells, FAB order
void ref_init_from_populations(void* h, const double* f0, const double* g0) {
  RefLattice* L = static_cast<RefLattice*>(h);
  MultiFab mf0, mg0;
  mf0.define(L->geom.Domain(), nvel, 2); mg0.define(L->geom.Domain(), nvel, 2);
  copy_in(mf0, nvel, f0); copy_in(mg0, nvel, g0);
  LBM_init(L->geom, L->fold, L->gold, L->hydrovs, L->hydrovsbar, L->fnoise, L->gnoise, mf0, mg0, L->rho_eq, L->phi_eq, L->rhot_eq, L->com_ref);
}
// The fluctuating run's equilibrium profiles (main_run_job.cpp:216-236: LoadSingleMultiFab of equilibrium_{rho,phi,rhot},
// com_ref = their centres of mass).  They only matter in the build with -DUSE_REF_STATE (LBM_binary.H:12, 92-107).
void ref_set_equilibrium(void* h, const double* rho_eq, const double* phi_eq, const double* rhot_eq) {
  RefLattice* L = static_cast<RefLattice*>(h);
  copy_in(L->rho_eq, 1, rho_eq); copy_in(L->phi_eq, 1, phi_eq); copy_in(L->rhot_eq, 1, rhot_eq);
  RealVect com_rho, com_phi, com_rhot;
  update_com(L->geom, com_rho, L->rho_eq, true);
  update_com(L->geom, com_phi, L->phi_eq, true);
  update_com(L->geom, com_rhot, L->rhot_eq, true);
  L->com_ref[0] = com_rho; L->com_ref[1] = com_phi; L->com_ref[2] = com_rhot;
}
int ref_uses_ref_state() {
#ifdef USE_REF_STATE
  return 1;
#else
  return 0;
#endif
}
void ref_step(void* h, int nsteps) {
  RefLattice* L = static_cast<RefLattice*>(h);
  for (int s = 0; s < nsteps; ++s)
    LBM_timestep(L->geom, L->fold, L->gold, L->fnew, L->gnew, L->hydrovs, L->hydrovsbar, L->fnoise, L->gnoise, L->rho_eq, L->phi_eq, L->rhot_eq, L->com_ref);
}
// valid cells, FAB order (x fastest ... component slowest)
void ref_get_populations(void* h, double* f, double* g) {
  RefLattice* L = static_cast<RefLattice*>(h);
  copy_out(L->fold, 0, nvel, f); copy_out(L->gold, 0, nvel, g);
}
void ref_get_hydrovars(void* h, double* out22) { copy_out(static_cast<RefLattice*>(h)->hydrovs, 0, 22, out22); }
void ref_get_hydrovars_bar(void* h, double* out9) { copy_out(static_cast<RefLattice*>(h)->hydrovsbar, 0, 9, out9); }
void ref_get_noise(void* h, double* fn, double* gn) {
  RefLattice* L = static_cast<RefLattice*>(h);
  copy_out(L->fnoise, 0, nvel, fn); copy_out(L->gnoise, 0, nvel, gn);
}
// lattice constants and the two transforms, for unit checks of the restatement
void ref_constants(int* c_out57, double* w_out19, double* b_out19) {
  for (int i = 0; i < nvel; ++i) {
    for (int d = 0; d < 3; ++d) c_out57[3 * i + d] = c[i][d];
    w_out19[i] = w[i]; b_out19[i] = b[i];
  }
}
void ref_moments(const double* f19, double* m19) {
  Array1D<Real, 0, nvel> f;
  for (int i = 0; i < nvel; ++i) f(i) = f19[i];
  Array1D<Real, 0, nvel> m = moments(f);
  for (int i = 0; i < nvel; ++i) m19[i] = m(i);
}
void ref_populations(const double* m19, double* f19) {
  Array1D<Real, 0, nvel> m;
  for (int i = 0; i < nvel; ++i) m(i) = m19[i];
  Array1D<Real, 0, nvel> f = populations(m);
  for (int i = 0; i < nvel; ++i) f19[i] = f(i);
}
int ref_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
// explicit thread count of the timing build (an OMP_NUM_THREADS=1 exported by a launcher must not decide the baseline)
void ref_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
}  // extern "C"
