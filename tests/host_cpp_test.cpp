// CPU test program for the host-side C++ (no GPU needed): Parameters parser and plotfile round trip.
#include <cassert>
#include <cmath>
#include <cstdio>
#include <sstream>

#include "../binary-fluctuating-lattice-boltzmann_b200/csrc/host/parameters.hpp"
#include "../binary-fluctuating-lattice-boltzmann_b200/csrc/host/plotfile.hpp"

using namespace bflbm;

#define EXPECT_THROW(stmt)                       \
  do {                                           \
    bool thrown = false;                         \
    try { stmt; } catch (const std::exception&) { thrown = true; } \
    assert(thrown && #stmt);                     \
  } while (0)

int main(int argc, char** argv) {
  const std::string tmp = argc > 1 ? argv[1] : "/tmp";
  {  // defaults = the values shipped in the reference's sources
    std::istringstream in("");
    RunParameters P = parse_parameters(in);
    assert(P.system == "mixture" && P.nx == 32 && P.ny == 32 && P.nz == 32);
    assert(P.nsteps == 40000 && P.plot_int == 200 && P.print_int == 20 && P.t_window == 1000 && P.out_noise_step == 40001);
    assert(P.kBT == 0. && P.tau_f == 0.5 && P.tau_g == 0.5 && P.alpha0 == 4. && P.kappa == 4. && P.rho_lo == 0. && P.rho_hi == 1. && P.seed == 12345ull);
    assert(P.out_step == 0 && P.radius == 0.2 && P.init_frac == 0.5 && P.Ndigits == 7);
  }
  {
    std::istringstream in("# comment\nsystem = flat_interface\nnx = 8; \nny=256\n nz = 64 // trailing\nkBT = 1e-5\nnsteps=800000\nstep_continue = 3000\nplot_int = 1000\nalpha0=1.5\nseed = 42\nif_continue_from_last_frame = true\n");
    RunParameters P = parse_parameters(in);
    assert(P.system == "flat_interface" && P.nx == 8 && P.ny == 256 && P.nz == 64 && P.kBT == 1e-5 && P.seed == 42 && P.if_continue_from_last_frame);
    assert(P.out_step == 3000 + 2 * 800000 / 10);  // main_run_job.cpp:89
  }
  { std::istringstream in("bogus_key = 1\n"); EXPECT_THROW(parse_parameters(in)); }
  { std::istringstream in("system = plasma\n"); EXPECT_THROW(parse_parameters(in)); }
  { std::istringstream in("use_SC_pseudo = true\n"); EXPECT_THROW(parse_parameters(in)); }
  { std::istringstream in("nsteps = 1001\nplot_int = 10\n"); EXPECT_THROW(parse_parameters(in)); }
  { std::istringstream in("nx 32\n"); EXPECT_THROW(parse_parameters(in)); }

  {  // plotfile round trip, 22 names, non-cubic box
    const int nx = 5, ny = 3, nz = 4, nc = 22;
    std::vector<double> d((size_t)nx * ny * nz * nc);
    for (size_t i = 0; i < d.size(); ++i) d[i] = std::sin(0.37 * (double)i) * 1e-3 + (double)(i % 7);
    std::vector<std::string> names;
    for (int c = 0; c < nc; ++c) names.push_back("v" + std::to_string(c));
    const std::string dir = concatenate(tmp + "/bflbm_test_plt", 200, 7);
    assert(dir.substr(dir.size() - 7) == "0000200");
    write_plotfile(dir, d, nc, nx, ny, nz, names, 200., 200);
    PlotfileData P = read_plotfile(dir);
    assert(P.ncomp == nc && P.nx == nx && P.ny == ny && P.nz == nz && P.time == 200. && P.names[21] == "v21");
    assert(P.data == d);
    // checkpoint style: 19 components under a single name (main_run_job.cpp:406-409)
    std::vector<double> f((size_t)nx * ny * nz * 19, 0.25);
    write_plotfile(tmp + "/bflbm_test_chk", f, 19, nx, ny, nz, {"rho_chk"}, 0, 0);
    PlotfileData C = read_plotfile(tmp + "/bflbm_test_chk");
    assert(C.ncomp == 19 && C.data == f && C.names[0] == "rho_chk");
    EXPECT_THROW(read_plotfile(tmp + "/does_not_exist"));
    // multi-box files (what the reference writes from its max_grid_size decomposition, main_run_job.cpp:140-143): boxes of
    // at most 2 cells per edge -> 3 x 2 x 2 = 12 FABs at increasing offsets, reassembled into the same global array
    write_plotfile(tmp + "/bflbm_test_multibox", d, nc, nx, ny, nz, names, 7., 7, 2);
    PlotfileData M = read_plotfile(tmp + "/bflbm_test_multibox");
    assert(M.nboxes == 12 && M.nx == nx && M.ny == ny && M.nz == nz && M.step == 7);
    assert(M.data == d);
  }
  {  // FrameWriter: jobs run in order on the writer thread, at most max_pending wait, errors come back to the caller
    std::vector<int> order;
    {
      FrameWriter w(2);
      for (int i = 0; i < 20; ++i) w.push([&order, i] { order.push_back(i); });
      w.drain();
      assert(order.size() == 20 && w.frames_written_async() == 20);
      for (int i = 0; i < 20; ++i) assert(order[i] == i);
      w.push([] { throw std::runtime_error("disk full"); });
      bool thrown = false;
      try { w.drain(); } catch (const std::runtime_error& e) { thrown = std::string(e.what()) == "disk full"; }
      assert(thrown);
      w.push([&order] { order.push_back(99); });  // usable again after the error was reported
      w.drain();
      assert(order.back() == 99);
    }
    FrameWriter inline_writer(0);  // async_output = 0: the job runs in the caller
    bool ran = false;
    inline_writer.push([&ran] { ran = true; });
    assert(ran && inline_writer.frames_written_async() == 0);
    // files written through the writer are the files written inline
    const int nx = 4, ny = 3, nz = 2;
    std::vector<double> d((size_t)nx * ny * nz * 9);
    for (size_t i = 0; i < d.size(); ++i) d[i] = 0.5 * (double)i;
    {
      FrameWriter w(1);
      for (int s = 0; s < 3; ++s) w.push([=] { write_plotfile(concatenate(tmp + "/bflbm_async_plt", s, 7), d, 9, nx, ny, nz, {"a"}, s, s); });
    }  // the destructor drains
    for (int s = 0; s < 3; ++s) assert(read_plotfile(concatenate(tmp + "/bflbm_async_plt", s, 7)).data == d);
    { std::istringstream in("async_output = 0\n"); assert(parse_parameters(in).async_output == 0); }
    { std::istringstream in("async_output = -1\n"); EXPECT_THROW(parse_parameters(in)); }
    { std::istringstream in(""); assert(parse_parameters(in).async_output == 2); }
  }
  std::puts("host_cpp_test ok");
  return 0;
}
