// Run-time parameter file for the driver.  The reference has none: every tunable is a global or a local
// edited in source (kBT LBM_d3q19.H:10; tau_f, tau_g, alpha0, alpha1, kappa, seed, rho_lo, rho_hi
// LBM_binary.H:17-30; system macro main_run_job.cpp:24-26; sizes, run lengths and output cadence
// main_run_job.cpp:71-103, 110-111, 28-33) and the shipped `Parameters` file is free-text notes of recipes.
// This parser reads "key = value" lines ('#' or '//' start a comment) with exactly those names.
#pragma once
#include <cctype>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>

namespace bflbm {

struct RunParameters {
  // main_run_job.cpp:24-26
  std::string system = "mixture";  // mixture | flat_interface | droplet
  // main_run_job.cpp:71-73, 124-134
  int nx = 32, ny = 0, nz = 0;  // 0 = same as nx
  int max_grid_size = 0;        // accepted and ignored (AMReX box size; the GPU layout is one slab per GPU)
  // main_run_job.cpp:80-103
  int step_continue = 0;
  bool continueFromNonFluct = true;
  bool if_continue_from_last_frame = false;
  int nsteps = 40000;
  int out_step = -1;  // -1 = reference default: kBT ? step_continue + 2*nsteps/10 : step_continue
  int plot_int = 200;
  int print_int = 20;
  int t_window = -1;         // -1 = 5*plot_int
  int out_noise_step = -1;   // -1 = nsteps+1 (never)
  int plot_SF_window = 0;
  int out_SF_step = 100;
  // main_run_job.cpp:110-111, 33, 28-29
  double radius = 0.2;
  bool if_print_radius = false;
  double init_frac = 0.5;
  std::string root_path = ".";
  int Ndigits = 7;
  std::string plot_fields = "hydrovars_bar";  // STRUCT_LB_HYDROVARS (shipped, main_run_job.cpp:19) | hydrovars (STRUCT_HYDROVARS)
  int device = 0;
  int ngpus = 1;     // z-slabs over devices device .. device+ngpus-1 of this process (the reference: mpirun ranks + max_grid_size)
  int brick_lz = 0;  // brick height of the step kernel; 0 = automatic.  Same value => same bits on any GPU count
  int async_output = 2;  // frames waiting for the writer thread (plotfiles are written while the GPUs run the next interval);
                         // 0 = write inline like the reference (main_run_job.cpp:372-385)
  // LBM_d3q19.H:10, LBM_binary.H:17-30
  double kBT = 0., tau_f = 0.5, tau_g = 0.5, alpha0 = 4., alpha1 = 0., kappa = 4., rho_lo = 0., rho_hi = 1.;
  unsigned long long seed = 12345ull;
  bool use_ref_state = false;  // the reference's USE_REF_STATE macro (LBM_binary.H:12, shipped commented out): noise amplitudes from
                               // the equilibrium_* profiles of the kBT = 0 run, shifted with the centre of mass
  bool use_SC_pseudo = false;  // dead branch in the reference (LBM_binary.H:23); true is rejected
  double SC_ref_density = 1.;
};

inline std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && std::isspace((unsigned char)s[a])) ++a;
  while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}
inline bool to_bool(const std::string& v, const std::string& key) {
  if (v == "true" || v == "1" || v == "TRUE" || v == "True") return true;
  if (v == "false" || v == "0" || v == "FALSE" || v == "False") return false;
  throw std::runtime_error("Parameters: bad boolean for " + key + ": " + v);
}

inline RunParameters parse_parameters(std::istream& in) {
  RunParameters P;
  std::map<std::string, std::string> kv;
  std::string line;
  int lineno = 0;
  while (std::getline(in, line)) {
    ++lineno;
    size_t c = line.find('#');
    if (c != std::string::npos) line = line.substr(0, c);
    c = line.find("//");
    if (c != std::string::npos) line = line.substr(0, c);
    line = trim(line);
    if (line.empty()) continue;
    size_t eq = line.find('=');
    if (eq == std::string::npos) throw std::runtime_error("Parameters: line " + std::to_string(lineno) + ": expected key = value");
    std::string key = trim(line.substr(0, eq)), val = trim(line.substr(eq + 1));
    if (!val.empty() && val.back() == ';') val = trim(val.substr(0, val.size() - 1));
    kv[key] = val;
  }
  auto take = [&](const char* k, auto setter) {
    auto it = kv.find(k);
    if (it == kv.end()) return;
    try {
      setter(it->second);
    } catch (const std::exception& e) {
      throw std::runtime_error(std::string("Parameters: bad value for ") + k + ": " + it->second);
    }
    kv.erase(it);
  };
#define P_INT(name) take(#name, [&](const std::string& v) { P.name = std::stoi(v); })
#define P_DBL(name) take(#name, [&](const std::string& v) { P.name = std::stod(v); })
#define P_BOOL(name) take(#name, [&](const std::string& v) { P.name = to_bool(v, #name); })
#define P_STR(name) take(#name, [&](const std::string& v) { P.name = v; })
  P_STR(system); P_INT(nx); P_INT(ny); P_INT(nz); P_INT(max_grid_size);
  P_INT(step_continue); P_BOOL(continueFromNonFluct); P_BOOL(if_continue_from_last_frame);
  P_INT(nsteps); P_INT(out_step); P_INT(plot_int); P_INT(print_int); P_INT(t_window); P_INT(out_noise_step);
  P_INT(plot_SF_window); P_INT(out_SF_step);
  P_DBL(radius); P_BOOL(if_print_radius); P_DBL(init_frac); P_STR(root_path); P_INT(Ndigits); P_STR(plot_fields); P_INT(device); P_INT(ngpus); P_INT(brick_lz); P_INT(async_output);
  P_DBL(kBT); P_DBL(tau_f); P_DBL(tau_g); P_DBL(alpha0); P_DBL(alpha1); P_DBL(kappa); P_DBL(rho_lo); P_DBL(rho_hi);
  take("seed", [&](const std::string& v) { P.seed = std::stoull(v); });
  P_BOOL(use_ref_state); P_BOOL(use_SC_pseudo); P_DBL(SC_ref_density);
#undef P_INT
#undef P_DBL
#undef P_BOOL
#undef P_STR
  if (!kv.empty()) throw std::runtime_error("Parameters: unknown key '" + kv.begin()->first + "'");
  if (P.system != "mixture" && P.system != "flat_interface" && P.system != "droplet")
    throw std::runtime_error("Parameters: system must be mixture | flat_interface | droplet");
  if (P.use_SC_pseudo) throw std::runtime_error("Parameters: use_SC_pseudo = true is a dead branch in the reference and is not supported");
  if (P.plot_fields != "hydrovars" && P.plot_fields != "hydrovars_bar") throw std::runtime_error("Parameters: plot_fields must be hydrovars | hydrovars_bar");
  if (P.ngpus < 1) throw std::runtime_error("Parameters: ngpus must be >= 1");
  if (P.async_output < 0) throw std::runtime_error("Parameters: async_output must be >= 0");
  if (P.ny <= 0) P.ny = P.nx;
  if (P.nz <= 0) P.nz = P.nx;
  if (P.t_window < 0) P.t_window = 5 * P.plot_int;
  if (P.out_noise_step < 0) P.out_noise_step = P.nsteps + 1;
  if (P.out_step < 0) P.out_step = (P.kBT != 0.) ? P.step_continue + 2 * P.nsteps / 10 : P.step_continue;
  if (P.plot_int > 0 && P.nsteps % P.plot_int != 0)
    throw std::runtime_error("Parameters: nsteps must be an integer multiple of plot_int (main_run_job.cpp:86)");
  return P;
}
inline RunParameters parse_parameters_file(const std::string& path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("cannot open parameter file " + path);
  return parse_parameters(in);
}

}  // namespace bflbm
