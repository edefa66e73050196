// C++ host-side mirror of the reference's solver interface over the C ABI (include/bflbm.h).
//
// The reference's driver calls header-inline functions on caller-owned MultiFabs:
//   LBM_init_mixture / LBM_init_stripe / LBM_init_droplet / LBM_init / LBM_timestep / LBM_hydrovars /
//   LBM_hydrovars_density / thermal_noise / update_com / fittingDropletCovariance   (LBM_binary.H:73-742, LBM_hydrovs.H:26-60, 258-335)
// plus globals kBT, tau_f, tau_g, alpha0, alpha1, kappa, seed (LBM_d3q19.H:10, LBM_binary.H:17-30).
// bflbm::Lattice bundles what those MultiFabs hold (device resident) and the free functions below keep the
// reference's names and argument meaning, so a driver written against LBM_binary.H ports line by line.
// Errors: the reference returns void and exit(0)s on NaN (Debug.H:137-149); here every failure throws bflbm::Error.
#pragma once
#include <array>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../../include/bflbm.h"

namespace bflbm {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
  if (rc != 0) throw Error(rc, std::string("bflbm: ") + bflbm_last_error());
}

// the reference's global parameters, with its shipped defaults
inline bflbm_params default_params() {
  bflbm_params p;
  check(bflbm_params_default(&p));
  return p;
}

inline void mcheck(int rc) {
  if (rc != 0) throw Error(rc, std::string("bflbm: ") + bflbm_multi_last_error());
}

// The MultiFab bundle of one periodic box.  ngpus > 1: the box is cut into z-slabs, one per device of this process, that
// exchange their ghost planes by peer-to-peer stores (bflbm_multi, include/bflbm.h) -- what BoxArray::maxSize +
// DistributionMapping + FillBoundary do in the reference (main_run_job.cpp:140-145).  Host arrays are always those of the
// whole box.  brick_lz: brick height of the step kernel (0 = automatic); runs on different GPU counts agree bit for bit
// when they use the same height and it divides every slab.
class Lattice {
 public:
  Lattice(const bflbm_params& p, int nx, int ny, int nz, int device = 0, int ngpus = 1, int brick_lz = 0) : nx_(nx), ny_(ny), nz_(nz) {
    std::vector<int> dev(ngpus);
    for (int i = 0; i < ngpus; ++i) dev[i] = device + i;
    mcheck(bflbm_multi_create(&p, nx, ny, nz, ngpus, dev.data(), brick_lz, &m_));
  }
  ~Lattice() { bflbm_multi_destroy(m_); }
  Lattice(const Lattice&) = delete;
  Lattice& operator=(const Lattice&) = delete;

  int nx() const { return nx_; }
  int ny() const { return ny_; }
  int nz() const { return nz_; }
  int ngpus() const { return bflbm_multi_count(m_); }
  size_t cells() const { return (size_t)nx_ * ny_ * nz_; }
  bflbm_multi* multi() const { return m_; }
  // the single lattice of a one-GPU box (structure-factor accumulator); null for ngpus > 1
  bflbm_lattice* handle() const { return ngpus() == 1 ? bflbm_multi_slab(m_, 0) : nullptr; }
  long long step_count() const { return bflbm_multi_step_count(m_); }
  void set_params(const bflbm_params& p) { mcheck(bflbm_multi_set_params(m_, &p)); }
  void sync() { mcheck(bflbm_multi_sync(m_)); }

  // FAB-order host arrays (x fastest ... component slowest), valid region
  std::vector<double> hydrovars() { std::vector<double> v(BFLBM_NHYDRO * cells()); mcheck(bflbm_multi_get_hydrovars(m_, v.data())); return v; }
  std::vector<double> hydrovars_bar() { std::vector<double> v(BFLBM_NHYDRO_BAR * cells()); mcheck(bflbm_multi_get_hydrovars_bar(m_, v.data())); return v; }
  std::pair<std::vector<double>, std::vector<double>> populations() {
    std::vector<double> f(BFLBM_NVEL * cells()), g(BFLBM_NVEL * cells());
    mcheck(bflbm_multi_get_populations(m_, f.data(), g.data()));
    return {std::move(f), std::move(g)};
  }
  std::pair<std::vector<double>, std::vector<double>> noise() {
    std::vector<double> f(BFLBM_NVEL * cells()), g(BFLBM_NVEL * cells());
    mcheck(bflbm_multi_get_noise(m_, f.data(), g.data()));
    return {std::move(f), std::move(g)};
  }

 private:
  bflbm_multi* m_ = nullptr;
  int nx_, ny_, nz_;
};

// ---- the reference's entry points (same names; the MultiFab bundle is the Lattice) -------------------
inline void LBM_init_mixture(Lattice& L) { mcheck(bflbm_multi_init_mixture(L.multi())); }                      // LBM_binary.H:598
inline void LBM_init_stripe(double frac, Lattice& L) { mcheck(bflbm_multi_init_stripe(L.multi(), frac)); }     // LBM_binary.H:664
inline void LBM_init_droplet(double r, Lattice& L) { mcheck(bflbm_multi_init_droplet(L.multi(), r)); }         // LBM_binary.H:699
inline void LBM_init(Lattice& L, const std::vector<double>& f0, const std::vector<double>& g0) {        // LBM_binary.H:632
  if (f0.size() != BFLBM_NVEL * L.cells() || g0.size() != f0.size()) throw Error(BFLBM_ERR_ARG, "LBM_init: population array size");
  mcheck(bflbm_multi_init_from_populations(L.multi(), f0.data(), g0.data()));
}
inline void LBM_timestep(Lattice& L, int nsteps = 1) { mcheck(bflbm_multi_step(L.multi(), nsteps)); }           // LBM_binary.H:545
inline std::vector<double> LBM_hydrovars(Lattice& L) { return L.hydrovars(); }                            // LBM_binary.H:298
inline std::vector<double> LBM_hydrovars_density(Lattice& L) { return L.hydrovars_bar(); }                // LBM_binary.H:343
inline std::pair<std::vector<double>, std::vector<double>> thermal_noise(Lattice& L) { return L.noise(); }  // LBM_binary.H:74
inline std::array<double, 3> update_com(Lattice& L) {                                                     // LBM_hydrovs.H:27
  std::array<double, 3> c;
  mcheck(bflbm_multi_center_of_mass(L.multi(), c.data()));
  return c;
}
// fittingDropletCovariance (LBM_hydrovs.H:258-335), one frame: eigenvalues (ascending) of the mass-weighted covariance of rho
inline std::array<double, 3> fittingDropletCovariance(Lattice& L) {
  std::array<double, 3> e;
  mcheck(bflbm_multi_droplet_covariance(L.multi(), nullptr, nullptr, e.data()));
  return e;
}
// MultiFabNANCheck (Debug.H:136-149): throws instead of exit(0)
inline void MultiFabNANCheck(Lattice& L) {
  long long n = 0;
  mcheck(bflbm_multi_check_nan(L.multi(), &n));
}

// hydrovs component names, AMReX_FileIO.H:208-261
inline std::vector<std::string> VariableNames(int n = BFLBM_NHYDRO) {
  static const char* names[BFLBM_NHYDRO] = {"rho", "phi", "ufx", "ufy", "ufz", "p_bulk", "ugx", "ugy", "ugz", "afx", "afy",
                                            "afz", "agx", "agy", "agz", "ubx", "uby", "ubz", "nfbarx", "ngbarx", "ufbarx", "ugbarx"};
  std::vector<std::string> v;
  for (int i = 0; i < n && i < BFLBM_NHYDRO; ++i) v.push_back(names[i]);
  return v;
}

}  // namespace bflbm
