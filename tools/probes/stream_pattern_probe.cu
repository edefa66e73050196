// Probe: how much HBM bandwidth does the step kernel's ACCESS PATTERN allow, with no arithmetic?
// 76 streams per CTA (38 components read with the D3Q19 pull shifts, 38 written), brick sweep over lz planes,
//   layout 0: linear rows (x fastest, row stride nx*8 B)          -- what the library uses
//   layout 1: blocked (one 32x8 tile plane = 2 KB contiguous per component)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_pattern_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
constexpr int Q = 19;
__host__ __device__ constexpr int cx(int i) { constexpr int v[Q] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1}; return v[i]; }
__host__ __device__ constexpr int cy(int i) { constexpr int v[Q] = {0, 0, 0, 1, -1, 0, 0, 1, -1, -1, 1, 1, -1, 1, -1, 0, 0, 0, 0}; return v[i]; }
__host__ __device__ constexpr int cz(int i) { constexpr int v[Q] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1}; return v[i]; }
struct Bases { const double* in[2 * Q]; double* out[2 * Q]; };

template <int LAYOUT, int WORK, int NBAR = 0>
__global__ void __launch_bounds__(256, 2) k(const __grid_constant__ Bases B, int nx, int ny, int nz, int lz) {
  const int tx = threadIdx.x, ty = threadIdx.y, x = blockIdx.x * 32 + tx, y = blockIdx.y * 8 + ty, zb = blockIdx.z * lz;
  const int bx = nx / 32;
  auto idx = [&](int xx, int yy) -> unsigned {  // in-plane element index, periodic
    xx = (xx + nx) % nx; yy = (yy + ny) % ny;
    if (LAYOUT == 0) return (unsigned)(yy * nx + xx);
    return (unsigned)((((yy >> 3) * bx + (xx >> 5)) << 8) + ((yy & 7) << 5) + (xx & 31));
  };
  const unsigned c0 = idx(x, y) * 8u;
  unsigned d[Q];
#pragma unroll
  for (int i = 0; i < Q; ++i) d[i] = idx(x - cx(i), y - cy(i)) * 8u;
  const unsigned pl8 = (unsigned)(nx * ny) * 8u;
  for (int k = 0; k < lz; ++k) {
    const unsigned zoff = (unsigned)(zb + k + 1) * pl8;
    double f[2 * Q];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int i = 0; i < Q; ++i)
        f[s * Q + i] = __ldg((const double*)((const char*)B.in[s * Q + i] + (zoff - (unsigned)cz(i) * pl8 + d[i])));
    if (WORK) {  // a dependent fp64 chain of WORK fmas per value, to mimic compute between load and store
#pragma unroll
      for (int j = 0; j < 2 * Q; ++j)
#pragma unroll
        for (int w = 0; w < WORK; ++w) f[j] = fma(f[j], 1.0000001, 1e-9);
    }
#pragma unroll
    for (int j = 0; j < 2 * Q; ++j) *(double*)((char*)B.out[j] + (zoff + c0)) = f[j];
    if (NBAR) {  // NBAR barrier-separated shared-memory read-modify-write phases per plane, like the density scatter
      __shared__ double2 acc[3 * 340];
#pragma unroll
      for (int b = 0; b < NBAR; ++b) {
        __syncthreads();
        double2 v = acc[(b % 3) * 340 + (ty + 1) * 34 + tx + 1];
        v.x += f[b]; v.y += f[b + 19];
        acc[(b % 3) * 340 + (ty + 1) * 34 + tx + 1] = v;
      }
    }
  }
}
template <int LAYOUT, int WORK, int NBAR = 0>
double run(const Bases& B, int n, int lz, int reps) {
  dim3 grid(n / 32, n / 8, n / lz), block(32, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 2; ++w) k<LAYOUT, WORK, NBAR><<<grid, block>>>(B, n, n, n, lz);
  cudaEventRecord(e0);
  for (int r = 0; r < reps; ++r) k<LAYOUT, WORK, NBAR><<<grid, block>>>(B, n, n, n, lz);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}
int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 384, lz = 32;
  const size_t comp = (size_t)(n + 2) * n * n;
  double *X, *Y;
  if (cudaMalloc(&X, 2 * Q * comp * 8) != cudaSuccess || cudaMalloc(&Y, 2 * Q * comp * 8) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(X, 0, 2 * Q * comp * 8);
  Bases B;
  for (int i = 0; i < 2 * Q; ++i) { B.in[i] = X + i * comp; B.out[i] = Y + i * comp; }
  const double bytes = 608.0 * n * n * (double)n;
  double t;
  t = run<0, 0>(B, n, lz, 5); printf("n=%d linear  layout, no work : %.3f ms  %.0f GB/s (608 B/cell)\n", n, t, bytes / t / 1e6);
  t = run<1, 0>(B, n, lz, 5); printf("n=%d blocked layout, no work : %.3f ms  %.0f GB/s\n", n, t, bytes / t / 1e6);
  t = run<0, 8>(B, n, lz, 5); printf("n=%d linear  layout, 8 fma/value: %.3f ms  %.0f GB/s\n", n, t, bytes / t / 1e6);
  t = run<1, 8>(B, n, lz, 5); printf("n=%d blocked layout, 8 fma/value: %.3f ms  %.0f GB/s\n", n, t, bytes / t / 1e6);
  t = run<0, 24>(B, n, lz, 5); printf("n=%d linear  layout, 24 fma/value: %.3f ms  %.0f GB/s\n", n, t, bytes / t / 1e6);
  t = run<1, 24>(B, n, lz, 5); printf("n=%d blocked layout, 24 fma/value: %.3f ms  %.0f GB/s\n", n, t, bytes / t / 1e6);
  t = run<0, 0, 4>(B, n, lz, 5); printf("n=%d linear, no work, 4 barriers/plane : %.3f ms  %.0f GB/s\n", n, t, bytes / t / 1e6);
  t = run<0, 24, 4>(B, n, lz, 5); printf("n=%d linear, 24 fma/value, 4 barriers/plane: %.3f ms  %.0f GB/s\n", n, t, bytes / t / 1e6);
  t = run<0, 24, 2>(B, n, lz, 5); printf("n=%d linear, 24 fma/value, 2 barriers/plane: %.3f ms  %.0f GB/s\n", n, t, bytes / t / 1e6);
  t = run<0, 48, 4>(B, n, lz, 5); printf("n=%d linear, 48 fma/value, 4 barriers/plane: %.3f ms  %.0f GB/s\n", n, t, bytes / t / 1e6);
  t = run<0, 48, 0>(B, n, lz, 5); printf("n=%d linear, 48 fma/value, no barrier: %.3f ms  %.0f GB/s\n", n, t, bytes / t / 1e6);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
