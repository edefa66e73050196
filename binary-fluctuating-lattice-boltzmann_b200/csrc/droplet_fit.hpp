// Droplet (W, R) fit: host side of SURVEY 8(f) row 4.
//
// The reference monitors a droplet run by fitting  rho(r) ~ 1/2 (1 + tanh((R - |r - r0|) / sqrt(2W)))  to the density of fluid f
// with a damped gradient flow in (W, R) (fittingDropletParams, LBM_hydrovs.H:160-213, called at main_run_job.cpp:358-369 behind
// if_print_radius).  One flow step (paramsVariations, externlib.H:368-403) needs
//   * two integrals over the lattice that involve the density field,  M_W = int rho (R - r') sech^2((R - r') / s) dV / s^3  and
//     M_R = int rho sech^2((R - r') / s) dV / s,  s = sqrt(2W), r' = |r - r0|, r0 = centre of mass  (externlib.H:246-340)
//     -- these are device reductions (k_fit_terms, kernels.cuh); the field never leaves the GPU;
//   * closed-form coefficients of the tanh profile itself (J.., K.., externlib.H:199-366), evaluated from truncated series
//     (externlib.H:21-197) -- scalar arithmetic, restated here.
// Coordinates are the reference's: unit cube, cell centres (i + 1/2) / n (main_run_job.cpp:137-138, LBM_hydrovs.H:92-96); the
// driver's MultiFab is cell centred, for which the reference's integral3D is the plain sum times the cell volume
// (AMReX_Analysis.H:458-500).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <functional>
#include <vector>

namespace bflbm {
namespace fit {

constexpr int SERIES_TERMS = 20;        // externlib.H:21 NumOfTerms
constexpr double MIN_LEN_SCALE = 1e-6;  // LBM_hydrovs.H:16

inline double binomial(int n, int k) {
  if (k > n) return 0.;
  double r = 1.;
  for (int i = 1; i <= k; ++i) r *= (double)(n - i + 1) / i;
  return r;
}

// Taylor coefficients of sech^4: sech x = sum_k A_k x^2k / (2k)! with the recursion A_k = - sum_{j<k} A_j C(2k, 2j) (Euler
// numbers), S_k = fourfold convolution of A_k / (2k)!  (getCoefS, externlib.H:54-86)
inline std::vector<double> sech4_coefficients(int n) {
  std::vector<double> a(n), ap(n), s(n);
  a[0] = 1.;
  for (int k = 1; k < n; ++k) {
    double acc = 0.;
    for (int j = 0; j < k; ++j) acc += a[j] * binomial(2 * k, 2 * j);
    a[k] = -acc;
  }
  for (int k = 0; k < n; ++k) {
    double fact = 1.;
    for (int i = 1; i <= 2 * k; ++i) fact *= i;
    ap[k] = a[k] / fact;
    double acc = 0.;
    for (int k1 = 0; k1 <= k; ++k1)
      for (int k2 = 0; k2 <= k - k1; ++k2)
        for (int k3 = 0; k3 <= k - k1 - k2; ++k3) acc += ap[k1] * ap[k2] * ap[k3] * ap[k - k1 - k2 - k3];
    s[k] = acc;
  }
  return s;
}

// int_0^inf x^n sech^4(x - c) dx, n = 2, 3, 4: the tails |x - c| > delta from sech^4 y = (16/6) sum_k (-1)^k (k+1)(k+2)(k+3)
// e^{-(2k+4)|y|} integrated by parts, the core |x - c| < delta from the Taylor series of sech^4 (integral_func2_series,
// externlib.H:102-158, with d = 1 and delta = 1 as every caller passes them).  long double like the reference.
inline double sech4_moment(int n, double c, const std::vector<double>& S) {
  const long double delta = 1.0L, r = c;
  long double total = 0.0L;
  long double nfact = 1.0L;
  for (int i = 2; i <= n; ++i) nfact *= i;
  for (int k = 0; k < (int)S.size(); ++k) {
    const long double e = 2 * k + 4, ie = 1.0L / e, decay = std::exp(-e * delta);
    // sum_m n!/(n-m)! (+-1)^m x^(n-m) / e^(m+1): the antiderivative of x^n e^{-+ e x}
    long double lower = 0.0L, upper = 0.0L, coef = 1.0L, iem = ie;
    for (int m = 0; m <= n; ++m) {
      lower += ((m & 1) ? -coef : coef) * iem * std::pow(r - delta, (long double)(n - m));
      upper += coef * iem * std::pow(r + delta, (long double)(n - m));
      coef *= (n - m);
      iem *= ie;
    }
    const long double origin = ((n & 1) ? 1.0L : -1.0L) * nfact * std::pow(ie, (long double)(n + 1)) * std::exp(-e * (long double)c);
    const long double tails = (lower * decay + origin) + upper * decay;
    const long double w = (16.0L / 6.0L) * (k + 1) * (k + 2) * (k + 3) * tails;
    total = (k & 1) ? total - w : total + w;
    long double core = 0.0L;
    for (int l = 0; l <= n; ++l) {
      const int p = 2 * k + l + 1;
      const long double odd = std::pow(delta, (long double)p) - std::pow(-delta, (long double)p);
      core += (long double)binomial(n, l) * std::pow(r, (long double)(n - l)) * odd / p;
    }
    total += (long double)S[k] * core;
  }
  return (double)total;
}

// integral_func3_series (externlib.H:160-174) and integral_func1_series (:177-197): alternating sums of e^{-2kc} / k^2, k^3
inline double series3(int n, double c, int terms = 50) {
  double v = 0.;
  for (int k = 1; k <= terms; ++k) {
    const double k2 = (double)k * k, sgn = (k & 1) ? 1. : -1.;  // (-1)^(k+1)
    if (n == 3) v += 6. * sgn * (c / k2 + 0.25 / (k2 * k) * std::exp(-2. * k * c));
    else v += -sgn * std::exp(-2. * k * c) / k2 + sgn * 2. / k2;
  }
  return v + 2. * std::pow(c, n);
}
inline double series1(int n, double a, int terms = 100) {
  if (n != 3) return -a - std::log(2.) - std::log(std::cosh(a));
  double s1 = 0., s2 = 0.;
  for (int k = 1; k <= terms; ++k) {
    const double k2 = (double)k * k, sgn = (k & 1) ? 1. : -1.;
    s1 += sgn / k2 * std::exp(-2. * k * a);
    s2 += sgn / k2;
  }
  return 1.5 * s1 - 3. * s2 - 3. * a * a;
}

struct Coefficients {
  double JRR, JWR, JRW, JWW;  // JRn_Rn, JWn_Rn, JRn_Wn, JWn_Wn  (externlib.H:203-244)
  double KW, KR;              // KWn, KRn                         (externlib.H:342-366)
};
inline Coefficients coefficients(double W, double R, double eta_W, double eta_R, double dt, double C0, const std::vector<double>& S) {
  const double s = std::sqrt(2. * W), c = R / s;
  const double I2 = sech4_moment(2, c, S), I3 = sech4_moment(3, c, S), I4 = sech4_moment(4, c, S);
  Coefficients q;
  q.JRR = -C0 * eta_R * dt * s * M_PI * I2;
  q.JRW = C0 * 0.25 * eta_R * dt * M_PI / (W * W) * (R * 2. * W * s * I2 - 4. * W * W * I3);
  q.JWR = C0 * 0.25 * eta_W * dt * (2. * std::sqrt(2.) * M_PI * R / std::sqrt(W) * I2 - 4. * M_PI * I3);
  q.JWW = -C0 * 0.125 * eta_W * dt * M_PI / std::pow(W, 3) * (std::pow(s, 3) * R * R * I2 + std::pow(s, 5) * I4 - 2. * R * std::pow(s, 4) * I3);
  const double a2 = series3(2, c), a3 = series3(3, c), b2 = series1(2, c), b3 = series1(3, c);
  q.KW = std::sqrt(2.) * M_PI / std::pow(std::sqrt(W), 3.) * (R * std::pow(s, 3.) * a2 - 4. * W * W * a3 + R * std::pow(s, 3.) * b2 - 4. * W * W * b3);
  q.KR = 4. * M_PI * 2. * W * (a2 + b2);
  return q;
}

// The lattice side: for given (W, R) return the two raw sums  sum rho (R - r') sech^2((R - r')/s)  and  sum rho sech^2(..)  over all
// cells (any number of GPUs), r' measured from r0 in unit-cube coordinates.
using FieldTerms = std::function<int(double W, double R, const double r0[3], double sums[2])>;

struct Result {
  double W = 0., R = 0., undulation = 0.;
  bool converged = false;
  int flows = 0;  // gradient flows run (1 + retries)
  std::vector<std::array<double, 2>> trace;  // (W, R) of the last flow
};

// one gradient flow, fittingDroplet (LBM_hydrovs.H:114-146); rho_range = max rho - min rho (C0)
inline int flow(const FieldTerms& terms, double cell_volume, const double r0[3], double rho_range, double W0, double R0, double eta_W, double eta_R,
                double dt, int nstep, const std::vector<double>& S, std::vector<std::array<double, 2>>& trace) {
  trace.assign(nstep, {W0, R0});
  double W = W0, R = R0;
  for (int k = 1; k < nstep; ++k) {
    const Coefficients q = coefficients(W, R, eta_W, eta_R, dt, rho_range, S);
    double sums[2];
    if (const int rc = terms(W, R, r0, sums)) return rc;
    const double s = std::sqrt(2. * W);
    const double MW = sums[0] * cell_volume / std::pow(s, 3.), MR = sums[1] * cell_volume / s;
    // paramsVariations (externlib.H:368-403): delta = A B c / det, A = [[1 - JRR, JWR], [JRW, 1 - JWW]], B = diag(-eta_W dt, eta_R dt)
    const double c0 = MW - 0.5 * q.KW, c1 = MR - 0.5 * q.KR;
    const double A[2][2] = {{1. - q.JRR, q.JWR}, {q.JRW, 1. - q.JWW}}, B[2] = {-eta_W * dt, eta_R * dt};
    const double det = (1. - q.JWW) * (1. - q.JRR) - q.JWR * q.JRW;
    const double dW = (A[0][0] * B[0] * c0 + A[0][1] * B[1] * c1) / det, dR = (A[1][0] * B[0] * c0 + A[1][1] * B[1] * c1) / det;
    W += dW;
    R += dR;
    if (W <= 0.) {  // too large a step: undo it for W and slow down (LBM_hydrovs.H:134-137)
      W -= dW;
      dt /= 5.;
    }
    if (std::fabs(W) < MIN_LEN_SCALE) W = W0;
    trace[k] = {W, R};
  }
  return 0;
}

// fittingDropletParams (LBM_hydrovs.H:160-213): mean of the last `window` flow steps; while their spread exceeds `undul_ratio`,
// restart from the mean with a five times smaller step, at most 10 times.  The reference throws when that fails; here
// Result::converged is false.
inline int fit(const FieldTerms& terms, double cell_volume, const double r0[3], double rho_range, int window, double undul_ratio, int nstep,
               double W0, double R0, double eta_W, double eta_R, double dt, Result& out) {
  const std::vector<double> S = sech4_coefficients(SERIES_TERMS);
  if (window < 1 || window > nstep) return -1;
  auto spread = [&](double mean[2], double u[2]) {
    for (int c = 0; c < 2; ++c) {
      double acc = 0., mx = out.trace[nstep - window][c], mn = mx;
      for (int i = nstep - window; i < nstep; ++i) {
        acc += out.trace[i][c];
        mx = std::max(mx, out.trace[i][c]);
        mn = std::min(mn, out.trace[i][c]);
      }
      mean[c] = acc / window;
      u[c] = (mx - mn) / mean[c];
    }
  };
  double mean[2], u[2];
  if (const int rc = flow(terms, cell_volume, r0, rho_range, W0, R0, eta_W, eta_R, dt, nstep, S, out.trace)) return rc;
  out.flows = 1;
  spread(mean, u);
  double dt_new = dt / 5.;
  for (int it = 1; it <= 10 && !(u[0] <= undul_ratio && u[1] <= undul_ratio); ++it) {
    if (const int rc = flow(terms, cell_volume, r0, rho_range, mean[0], mean[1], eta_W, eta_R, dt_new, nstep, S, out.trace)) return rc;
    ++out.flows;
    spread(mean, u);
    dt_new /= 5.;
  }
  out.W = mean[0];
  out.R = mean[1];
  out.undulation = std::max(u[0], u[1]);
  out.converged = u[0] <= undul_ratio && u[1] <= undul_ratio;
  return 0;
}

}  // namespace fit
}  // namespace bflbm
