// AMReX single-level plotfile writer / reader (Header + Level_0/Cell_H + Level_0/Cell_D_00000).
// Replaces WriteSingleLevelPlotfile (main_run_job.cpp:48, 407-409, 432-438) and LoadSingleMultiFab =
// VisMF::Read of "<plotfile>/Level_0/Cell" (AMReX_FileIO.H:18-34) for the files this driver writes.
// The on-disk layout is AMReX's (third-party, not in the reference tree): restated from the AMReX plotfile
// convention and validated here by round trip only (single- and multi-box); valid cells, native
// little-endian doubles, x fastest ... component slowest inside each FAB.
#pragma once
#include <sys/stat.h>

#include <algorithm>
#include <cerrno>
#include <condition_variable>
#include <deque>
#include <exception>
#include <functional>
#include <mutex>
#include <thread>
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
// max_grid_size > 0 cuts the domain into boxes of at most that edge like BoxArray::maxSize (main_run_job.cpp:141): one FAB
// per box, all in Cell_D_00000 at increasing offsets -- the shape of the files the reference itself writes.
inline void write_plotfile(const std::string& dir, const std::vector<double>& data, int ncomp, int nx, int ny, int nz,
                           std::vector<std::string> names, double time, int step, int max_grid_size = 0) {
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
  struct Bx { int lo[3], hi[3]; };
  std::vector<Bx> boxes;
  {
    const int dims[3] = {nx, ny, nz};
    int cnt[3], len[3];
    for (int d = 0; d < 3; ++d) {
      cnt[d] = max_grid_size > 0 ? (dims[d] + max_grid_size - 1) / max_grid_size : 1;
      len[d] = (dims[d] + cnt[d] - 1) / cnt[d];
    }
    for (int k = 0; k < cnt[2]; ++k)
      for (int j = 0; j < cnt[1]; ++j)
        for (int i = 0; i < cnt[0]; ++i) {
          const int ijk[3] = {i, j, k};
          Bx b;
          for (int d = 0; d < 3; ++d) { b.lo[d] = ijk[d] * len[d]; b.hi[d] = std::min(dims[d], b.lo[d] + len[d]) - 1; }
          if (b.hi[0] >= b.lo[0] && b.hi[1] >= b.lo[1] && b.hi[2] >= b.lo[2]) boxes.push_back(b);
        }
  }
  auto box_str = [](const Bx& b) {
    std::ostringstream o;
    o << "((" << b.lo[0] << ',' << b.lo[1] << ',' << b.lo[2] << ") (" << b.hi[0] << ',' << b.hi[1] << ',' << b.hi[2] << ") (0,0,0))";
    return o.str();
  };
  std::vector<long long> offsets;
  std::vector<std::vector<double>> mins(boxes.size(), std::vector<double>(ncomp)), maxs = mins;
  {
    std::ofstream d(dir + "/Level_0/Cell_D_00000", std::ios::binary);
    std::vector<double> buf;
    for (size_t q = 0; q < boxes.size(); ++q) {
      const Bx& b = boxes[q];
      offsets.push_back((long long)d.tellp());
      d << "FAB ((8, (64 11 52 0 1 12 0 1023)),(8, (8 7 6 5 4 3 2 1)))" << box_str(b) << ' ' << ncomp << '\n';
      const int bx = b.hi[0] - b.lo[0] + 1, by = b.hi[1] - b.lo[1] + 1, bz = b.hi[2] - b.lo[2] + 1;
      buf.resize((size_t)bx * by * bz * ncomp);
      size_t o = 0;
      for (int c = 0; c < ncomp; ++c) {
        double mn = std::numeric_limits<double>::infinity(), mx = -mn;
        for (int z = b.lo[2]; z <= b.hi[2]; ++z)
          for (int y = b.lo[1]; y <= b.hi[1]; ++y) {
            const double* src = data.data() + (size_t)c * n + ((size_t)z * ny + y) * nx + b.lo[0];
            for (int x = 0; x < bx; ++x) { buf[o++] = src[x]; mn = std::min(mn, src[x]); mx = std::max(mx, src[x]); }
          }
        mins[q][c] = mn; maxs[q][c] = mx;
      }
      d.write(reinterpret_cast<const char*>(buf.data()), (std::streamsize)(buf.size() * sizeof(double)));
    }
    if (!d) throw std::runtime_error("cannot write " + dir + "/Level_0/Cell_D_00000");
  }
  {
    std::ofstream c(dir + "/Level_0/Cell_H");
    c << std::setprecision(17);
    c << 1 << '\n' << 1 << '\n' << ncomp << '\n' << 0 << '\n';
    c << '(' << boxes.size() << " 0\n";
    for (const Bx& b : boxes) c << box_str(b) << '\n';
    c << ")\n" << boxes.size() << '\n';
    for (size_t q = 0; q < boxes.size(); ++q) c << "FabOnDisk: Cell_D_00000 " << offsets[q] << '\n';
    c << '\n';
    for (int pass = 0; pass < 2; ++pass) {
      c << boxes.size() << ',' << ncomp << '\n';
      for (size_t q = 0; q < boxes.size(); ++q) {
        for (int k = 0; k < ncomp; ++k) c << (pass == 0 ? mins[q][k] : maxs[q][k]) << ',';
        c << '\n';
      }
      c << '\n';
    }
    if (!c) throw std::runtime_error("cannot write " + dir + "/Level_0/Cell_H");
  }
}

struct PlotfileData {
  int ncomp = 0, nx = 0, ny = 0, nz = 0, step = 0, nboxes = 0;
  double time = 0.;
  std::vector<std::string> names;
  std::vector<double> data;
};

// Reads a single-level plotfile: the domain comes from Header, every "FabOnDisk: <file> <offset>" entry of
// Level_0/Cell_H is one FAB (own box in its header line) that is scattered into the global [ncomp][nz][ny][nx] array --
// the multi-box files the reference writes from its max_grid_size decomposition as well as this driver's single-box ones.
inline PlotfileData read_plotfile(const std::string& dir) {
  PlotfileData P;
  int dlo[3] = {0, 0, 0}, dhi[3] = {-1, -1, -1};
  {
    std::ifstream h(dir + "/Header");
    if (!h) throw std::runtime_error("cannot open " + dir + "/Header");
    std::string line;
    std::getline(h, line);
    h >> P.ncomp;
    if (!h || P.ncomp < 1) throw std::runtime_error(dir + ": bad Header");
    P.names.resize(P.ncomp);
    for (auto& s : P.names) h >> s;
    int dim, finest;
    h >> dim >> P.time >> finest;
    if (dim != 3 || finest != 0) throw std::runtime_error(dir + ": only single-level 3-D plotfiles are supported");
    std::string rest((std::istreambuf_iterator<char>(h)), std::istreambuf_iterator<char>());
    const size_t b = rest.find("((");
    if (b == std::string::npos || std::sscanf(rest.c_str() + b, "((%d,%d,%d) (%d,%d,%d)", &dlo[0], &dlo[1], &dlo[2], &dhi[0], &dhi[1], &dhi[2]) != 6)
      throw std::runtime_error(dir + ": cannot parse the domain box in Header");
    std::istringstream after(rest.substr(rest.find('\n', b) + 1));
    after >> P.step;
  }
  P.nx = dhi[0] - dlo[0] + 1; P.ny = dhi[1] - dlo[1] + 1; P.nz = dhi[2] - dlo[2] + 1;
  if (P.nx < 1 || P.ny < 1 || P.nz < 1) throw std::runtime_error(dir + ": empty domain");
  const size_t n = (size_t)P.nx * P.ny * P.nz;
  P.data.assign(n * P.ncomp, std::numeric_limits<double>::quiet_NaN());
  std::ifstream c(dir + "/Level_0/Cell_H");
  if (!c) throw std::runtime_error("cannot open " + dir + "/Level_0/Cell_H");
  std::string all((std::istreambuf_iterator<char>(c)), std::istreambuf_iterator<char>());
  size_t pos = all.find("FabOnDisk:"), covered = 0;
  if (pos == std::string::npos) throw std::runtime_error(dir + ": no FabOnDisk entry");
  std::vector<double> buf;
  for (; pos != std::string::npos; pos = all.find("FabOnDisk:", pos + 1)) {
    std::istringstream fod(all.substr(pos + 10, 256));
    std::string fname;
    long long offset = 0;
    fod >> fname >> offset;
    std::ifstream d(dir + "/Level_0/" + fname, std::ios::binary);
    if (!d) throw std::runtime_error("cannot open " + dir + "/Level_0/" + fname);
    d.seekg(offset);
    std::string fh;
    std::getline(d, fh);
    // "FAB ((8, (...)),(8, (...)))((lox,loy,loz) (hix,hiy,hiz) (0,0,0)) ncomp"
    size_t b = fh.find(")))((");
    b = b == std::string::npos ? fh.find("))((") : b;
    int lo[3], hi[3];
    if (b == std::string::npos || std::sscanf(fh.c_str() + fh.find("((", b + 2), "((%d,%d,%d) (%d,%d,%d)", &lo[0], &lo[1], &lo[2], &hi[0], &hi[1], &hi[2]) != 6)
      throw std::runtime_error(dir + ": cannot parse FAB header: " + fh);
    if (fh.find("(8, (8 7 6 5 4 3 2 1))") == std::string::npos) throw std::runtime_error(dir + ": FAB is not native little-endian float64: " + fh);
    const int nc = std::atoi(fh.substr(fh.rfind(' ') + 1).c_str());
    if (nc != P.ncomp) throw std::runtime_error(dir + ": component count mismatch between Header and FAB");
    // a FAB may carry ghost cells: keep the part inside the domain
    const int bx = hi[0] - lo[0] + 1, by = hi[1] - lo[1] + 1, bz = hi[2] - lo[2] + 1;
    if (bx < 1 || by < 1 || bz < 1) throw std::runtime_error(dir + ": empty FAB box");
    buf.resize((size_t)bx * by * bz * nc);
    d.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)(buf.size() * sizeof(double)));
    if (!d) throw std::runtime_error(dir + ": short read of FAB data");
    for (int k = 0; k < nc; ++k)
      for (int z = std::max(lo[2], dlo[2]); z <= std::min(hi[2], dhi[2]); ++z)
        for (int y = std::max(lo[1], dlo[1]); y <= std::min(hi[1], dhi[1]); ++y) {
          const int x0 = std::max(lo[0], dlo[0]), x1 = std::min(hi[0], dhi[0]);
          if (x1 < x0) continue;
          const double* src = buf.data() + (((size_t)k * bz + (z - lo[2])) * by + (y - lo[1])) * bx + (x0 - lo[0]);
          double* dst = P.data.data() + (size_t)k * n + ((size_t)(z - dlo[2]) * P.ny + (y - dlo[1])) * P.nx + (x0 - dlo[0]);
          std::memcpy(dst, src, (size_t)(x1 - x0 + 1) * sizeof(double));
          if (k == 0) covered += (size_t)(x1 - x0 + 1);
        }
    ++P.nboxes;
  }
  if (covered < n) throw std::runtime_error(dir + ": the FABs do not cover the domain");
  return P;
}

// Output off the critical path.  The reference writes every frame inline between two time steps (WriteOutput,
// main_run_job.cpp:372-385); at the reference's own job sizes a frame costs more host time than the steps between two
// frames cost GPU time.  Here the driver hands the downloaded frame to ONE writer thread and goes on queueing steps; at most
// `max_pending` frames wait (that bounds the host memory), push() blocks beyond that.  max_pending = 0: write inline.
// drain() returns when everything is on disk and rethrows the first error of the writer.
class FrameWriter {
 public:
  explicit FrameWriter(int max_pending) : max_pending_(max_pending) {
    if (max_pending_ > 0) worker_ = std::thread([this] { run(); });
  }
  ~FrameWriter() {
    try { drain(); } catch (...) {}
    if (worker_.joinable()) {
      { std::lock_guard<std::mutex> g(m_); stop_ = true; }
      cv_.notify_all();
      worker_.join();
    }
  }
  FrameWriter(const FrameWriter&) = delete;
  FrameWriter& operator=(const FrameWriter&) = delete;

  void push(std::function<void()> job) {
    if (max_pending_ <= 0) { job(); return; }
    std::unique_lock<std::mutex> g(m_);
    cv_.wait(g, [&] { return (int)q_.size() < max_pending_ || err_; });
    if (err_) { auto e = err_; err_ = nullptr; std::rethrow_exception(e); }
    q_.push_back(std::move(job));
    ++written_async_;
    g.unlock();
    cv_.notify_all();
  }
  void drain() {
    if (max_pending_ <= 0) return;
    std::unique_lock<std::mutex> g(m_);
    cv_.wait(g, [&] { return (q_.empty() && !busy_) || err_; });
    if (err_) { auto e = err_; err_ = nullptr; std::rethrow_exception(e); }
  }
  long long frames_written_async() const { return written_async_; }

 private:
  void run() {
    for (;;) {
      std::function<void()> job;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return stop_ || !q_.empty(); });
        if (q_.empty()) return;
        job = std::move(q_.front());
        q_.pop_front();
        busy_ = true;
      }
      std::exception_ptr e;
      try { job(); } catch (...) { e = std::current_exception(); }
      {
        std::lock_guard<std::mutex> g(m_);
        busy_ = false;
        if (e && !err_) err_ = e;
      }
      cv_.notify_all();
    }
  }
  int max_pending_;
  std::thread worker_;
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<std::function<void()>> q_;
  bool stop_ = false, busy_ = false;
  std::exception_ptr err_;
  long long written_async_ = 0;
};

}  // namespace bflbm
