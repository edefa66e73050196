// AMReX single-level plotfile writer / reader (Header + Level_0/Cell_H + Level_0/Cell_D_00000).
// Replaces WriteSingleLevelPlotfile (main_run_job.cpp:48, 407-409, 432-438) and LoadSingleMultiFab =
// VisMF::Read of "<plotfile>/Level_0/Cell" (AMReX_FileIO.H:18-34) for the files this driver writes.
// The on-disk layout is AMReX's (third-party, not in the reference tree): restated from the AMReX plotfile
// convention and validated here by round trip only; one box, valid cells, native
// little-endian doubles, x fastest ... component slowest.
#pragma once
#include <sys/stat.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace bflbm {

inline void make_dirs(const std::string& path) {
  std::string cur;
  for (size_t i = 0; i <= path.size(); ++i) {
    if (i == path.size() || path[i] == '/') {
      if (!cur.empty() && cur != ".") {
        if (mkdir(cur.c_str(), 0755) != 0 && errno != EEXIST) throw std::runtime_error("cannot create directory " + cur);
      }
    }
    if (i < path.size()) cur += path[i];
  }
}

// amrex::Concatenate(root, step, ndigits)
inline std::string concatenate(const std::string& root, int num, int mindigits) {
  std::ostringstream os;
  os << root << std::setw(mindigits) << std::setfill('0') << num;
  return os.str();
}

// data: [ncomp][nz][ny][nx].  names may be shorter than ncomp (the reference writes its 19-component
// checkpoints with a single name, main_run_job.cpp:406-409): missing names are filled with "<last>_<i>".
inline void write_plotfile(const std::string& dir, const std::vector<double>& data, int ncomp, int nx, int ny, int nz,
                           std::vector<std::string> names, double time, int step) {
  const size_t n = (size_t)nx * ny * nz;
  if (data.size() != n * ncomp) throw std::runtime_error("write_plotfile: data size mismatch");
  while ((int)names.size() < ncomp) names.push_back((names.empty() ? std::string("comp") : names.back()) + "_" + std::to_string(names.size()));
  make_dirs(dir + "/Level_0");
  {
    std::ofstream h(dir + "/Header");
    h << std::setprecision(17);
    h << "HyperCLaw-V1.1\n" << ncomp << '\n';
    for (int c = 0; c < ncomp; ++c) h << names[c] << '\n';
    h << 3 << '\n' << time << '\n' << 0 << '\n';
    h << "0 0 0\n1 1 1\n\n";
    h << "((0,0,0) (" << nx - 1 << ',' << ny - 1 << ',' << nz - 1 << ") (0,0,0))\n";
    h << step << '\n';
    h << 1.0 / nx << ' ' << 1.0 / ny << ' ' << 1.0 / nz << '\n';
    h << 0 << '\n' << 0 << '\n';
    h << "0 1 " << time << '\n' << step << '\n';
    h << "0 1\n0 1\n0 1\n";
    h << "Level_0/Cell\n";
    if (!h) throw std::runtime_error("cannot write " + dir + "/Header");
  }
  std::ostringstream fabhdr;
  fabhdr << "FAB ((8, (64 11 52 0 1 12 0 1023)),(8, (8 7 6 5 4 3 2 1)))((0,0,0) (" << nx - 1 << ',' << ny - 1 << ',' << nz - 1 << ") (0,0,0)) " << ncomp << '\n';
  {
    std::ofstream d(dir + "/Level_0/Cell_D_00000", std::ios::binary);
    d << fabhdr.str();
    d.write(reinterpret_cast<const char*>(data.data()), (std::streamsize)(data.size() * sizeof(double)));
    if (!d) throw std::runtime_error("cannot write " + dir + "/Level_0/Cell_D_00000");
  }
  {
    std::ofstream c(dir + "/Level_0/Cell_H");
    c << std::setprecision(17);
    c << 1 << '\n' << 1 << '\n' << ncomp << '\n' << 0 << '\n';
    c << "(1 0\n((0,0,0) (" << nx - 1 << ',' << ny - 1 << ',' << nz - 1 << ") (0,0,0))\n)\n";
    c << 1 << '\n' << "FabOnDisk: Cell_D_00000 0\n\n";
    for (int pass = 0; pass < 2; ++pass) {
      c << "1," << ncomp << '\n';
      for (int k = 0; k < ncomp; ++k) {
        auto b = data.begin() + (std::ptrdiff_t)(k * n), e = b + (std::ptrdiff_t)n;
        c << (pass == 0 ? *std::min_element(b, e) : *std::max_element(b, e)) << ',';
      }
      c << "\n\n";
    }
    if (!c) throw std::runtime_error("cannot write " + dir + "/Level_0/Cell_H");
  }
}

struct PlotfileData {
  int ncomp = 0, nx = 0, ny = 0, nz = 0, step = 0;
  double time = 0.;
  std::vector<std::string> names;
  std::vector<double> data;
};

// reads a plotfile written by write_plotfile (single box, FabOnDisk offset 0 or any offset)
inline PlotfileData read_plotfile(const std::string& dir) {
  PlotfileData P;
  {
    std::ifstream h(dir + "/Header");
    if (!h) throw std::runtime_error("cannot open " + dir + "/Header");
    std::string line;
    std::getline(h, line);
    h >> P.ncomp;
    P.names.resize(P.ncomp);
    for (auto& s : P.names) h >> s;
    int dim, finest;
    h >> dim >> P.time >> finest;
    if (dim != 3 || finest != 0) throw std::runtime_error(dir + ": only single-level 3-D plotfiles are supported");
  }
  std::ifstream c(dir + "/Level_0/Cell_H");
  if (!c) throw std::runtime_error("cannot open " + dir + "/Level_0/Cell_H");
  std::string all((std::istreambuf_iterator<char>(c)), std::istreambuf_iterator<char>());
  size_t pos = all.find("FabOnDisk:");
  if (pos == std::string::npos) throw std::runtime_error(dir + ": no FabOnDisk entry");
  if (all.find("FabOnDisk:", pos + 1) != std::string::npos) throw std::runtime_error(dir + ": multi-box plotfiles are not supported by this reader");
  std::istringstream fod(all.substr(pos + 10));
  std::string fname;
  long long offset = 0;
  fod >> fname >> offset;
  std::ifstream d(dir + "/Level_0/" + fname, std::ios::binary);
  if (!d) throw std::runtime_error("cannot open " + dir + "/Level_0/" + fname);
  d.seekg(offset);
  std::string fh;
  std::getline(d, fh);
  // "... ((lox,loy,loz) (hix,hiy,hiz) (0,0,0)) ncomp"
  size_t b = fh.find(")))((");
  if (b == std::string::npos) b = fh.find("))((");
  int lo[3], hi[3], nc = 0;
  if (b == std::string::npos || std::sscanf(fh.c_str() + fh.find("((", b + 2), "((%d,%d,%d) (%d,%d,%d)", &lo[0], &lo[1], &lo[2], &hi[0], &hi[1], &hi[2]) != 6)
    throw std::runtime_error(dir + ": cannot parse FAB header: " + fh);
  nc = std::atoi(fh.substr(fh.rfind(' ') + 1).c_str());
  P.nx = hi[0] - lo[0] + 1; P.ny = hi[1] - lo[1] + 1; P.nz = hi[2] - lo[2] + 1;
  if (nc != P.ncomp) throw std::runtime_error(dir + ": component count mismatch between Header and FAB");
  P.data.resize((size_t)P.nx * P.ny * P.nz * nc);
  d.read(reinterpret_cast<char*>(P.data.data()), (std::streamsize)(P.data.size() * sizeof(double)));
  if (!d) throw std::runtime_error(dir + ": short read of FAB data");
  return P;
}

}  // namespace bflbm
