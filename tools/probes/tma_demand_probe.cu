// Probe: TMA as the DEMAND path of the step kernel's access pattern, not as a prefetcher.
// Same traffic as stream_pattern_probe.cu (38 components pulled with the D3Q19 shifts, 38 written, brick sweep over lz planes,
// no arithmetic), but every tile moves through the tensor-memory accelerator:
//   load : cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes   (SASS UTMALDG), one 32 x 8 x 1 x 1 box of
//          f64 per component and plane, start coordinate = tile origin - c_i.  Only INTERIOR tiles are swept (EDGE = 0): a box that
//          starts at a negative coordinate raised "illegal instruction" on this driver (580.159) -- a periodic lattice would
//          need its wrap columns handled by a separate path anyway; irrelevant for a bandwidth probe
//   store: cp.async.bulk.tensor.4d.global.shared::cta.bulk_group                              (SASS UTMASTG) from the landed tile
// STAGES-deep ring of 38 x 2.25 KB = 85.5 KB per stage, one elected thread per CTA issues everything, mbarrier per stage.
// Variants: TMA load + TMA store (pure async copy: the upper bound of what TMA can give this pattern), TMA load + thread stores
// (threads read their cell from shared memory and st.global it: what a real kernel would do after the collision).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tma_demand_probe tma_demand_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

constexpr int Q = 19, NC = 2 * Q, TX = 32, TY = 8;
constexpr int TILE_BYTES = TX * TY * 8;  // 2 KB
// A box whose first row does not start on a 16-byte boundary faults ("illegal instruction", measured: c_x = +-1 shifts of 8-byte
// elements).  The x-shifted components are therefore fetched as 36-wide boxes that start at x0 - 2 (16-byte aligned) and contain
// both shifted windows; a thread reads its value at column tx + 2 - c_x.  +12.5 % shared-memory bytes on 20 of 38 tiles.
constexpr int WX = TX + 4, WIDE_BYTES = WX * TY * 8;  // 2304 B
constexpr int SLOT_BYTES = WIDE_BYTES;                // every component gets a wide slot (keeps the offsets uniform)
__host__ __device__ constexpr int cx(int i) { constexpr int v[Q] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1}; return v[i]; }
__host__ __device__ constexpr int cy(int i) { constexpr int v[Q] = {0, 0, 0, 1, -1, 0, 0, 1, -1, -1, 1, 1, -1, 1, -1, 0, 0, 0, 0}; return v[i]; }
__host__ __device__ constexpr int cz(int i) { constexpr int v[Q] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1}; return v[i]; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load4(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z, int c) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z), "r"(c) : "memory");
}
__device__ __forceinline__ void tma_store4(const CUtensorMap* map, const void* src, int x, int y, int z, int c) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(x), "r"(y), "r"(z), "r"(c) : "memory");
}

// MODE 0: TMA load + TMA store.  MODE 1: TMA load + per-thread st.global (all 256 threads).
constexpr int EDGE_SKIP = 1;  // tiles skipped at every in-plane edge of the box
template <int MODE, int STAGES>
__global__ void __launch_bounds__(256, 1) k_tma(const __grid_constant__ CUtensorMap in, const __grid_constant__ CUtensorMap inw,
                                                const __grid_constant__ CUtensorMap out, double* Y,
                                                int nx, int ny, int nz, int lz, int ncl, int shift) {
  extern __shared__ unsigned char smem[];
  // TMA wants its shared-memory boxes 128-byte aligned; dynamic shared memory starts behind the static barriers
  double* tiles = reinterpret_cast<double*>(smem + ((128u - (smem_u32(smem) & 127u)) & 127u));  // [STAGES][NC][TY][TX]
  __shared__ uint64_t full[STAGES];
  const int tid = threadIdx.y * TX + threadIdx.x;
  const int x0 = (blockIdx.x + EDGE_SKIP) * TX, y0 = (blockIdx.y + EDGE_SKIP) * TY, zb = blockIdx.z * lz;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue_loads = [&](int k) {  // plane zb + k into stage k % STAGES (elected thread)
    const int s = k % STAGES;
    uint32_t bytes = 0;
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (c < ncl) bytes += (shift && cx(c % Q) != 0) ? WIDE_BYTES : TILE_BYTES;
    mbar_expect_tx(&full[s], bytes);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int i = c % Q;
      if (c < ncl) {
        double* dst = tiles + ((size_t)s * NC + c) * (SLOT_BYTES / 8);
        if (shift && cx(i) != 0) tma_load4(dst, &inw, &full[s], x0 - 2, y0 - cy(i), zb + k + 1 - cz(i), c);
        else                     tma_load4(dst, &in, &full[s], x0, y0 - shift * cy(i), zb + k + 1 - shift * cz(i), c);
      }
    }
  };
  if (tid == 0)
    for (int k = 0; k < STAGES && k < lz; ++k) issue_loads(k);
  const size_t comp = (size_t)(nz + 2) * nx * ny;
  for (int k = 0; k < lz; ++k) {
    const int s = k % STAGES;
    const uint32_t parity = (uint32_t)(k / STAGES) & 1u;
    if (MODE == 0) {
      if (tid == 0) {
        mbar_wait(&full[s], parity);
#pragma unroll
        for (int c = 0; c < NC; ++c)
          if (c < ncl && !(shift && cx(c % Q) != 0)) tma_store4(&out, tiles + ((size_t)s * NC + c) * (SLOT_BYTES / 8), x0, y0, zb + k + 1, c);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the stage's tiles have been read: it may be refilled
        if (k + STAGES < lz) issue_loads(k + STAGES);
      }
    } else {
      mbar_wait(&full[s], parity);  // every thread observes the phase
      const size_t cell = ((size_t)(zb + k + 1) * ny + (y0 + threadIdx.y)) * nx + x0 + threadIdx.x;
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c < ncl) {
          const double* t = tiles + ((size_t)s * NC + c) * (SLOT_BYTES / 8);
          const int i = c % Q;
          Y[c * comp + cell] = (shift && cx(i) != 0) ? t[threadIdx.y * WX + threadIdx.x + 2 - cx(i)] : t[tid];
        }
      __syncthreads();  // all threads are done with the stage
      if (tid == 0 && k + STAGES < lz) issue_loads(k + STAGES);
    }
  }
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE, int STAGES>
double run(const CUtensorMap& in, const CUtensorMap& inw, const CUtensorMap& out, double* Y, int n, int lz, int reps, int ncl, int shift) {
  const dim3 grid(n / TX - 2 * EDGE_SKIP, n / TY - 2 * EDGE_SKIP, n / lz), block(TX, TY);
  const size_t smem = (size_t)STAGES * NC * SLOT_BYTES + 128;
  cudaFuncSetAttribute(k_tma<MODE, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 2; ++w) k_tma<MODE, STAGES><<<grid, block, smem>>>(in, inw, out, Y, n, n, n, lz, ncl, shift);
  cudaEventRecord(e0);
  for (int r = 0; r < reps; ++r) k_tma<MODE, STAGES><<<grid, block, smem>>>(in, inw, out, Y, n, n, n, lz, ncl, shift);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("MODE %d: kernel error: %s\n", MODE, cudaGetErrorString(e)); exit(1); }
  return ms / reps;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 384, lz = 32;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const size_t comp = (size_t)(n + 2) * n * n;
  double *X, *Y;
  if (cudaMalloc(&X, NC * comp * 8) != cudaSuccess || cudaMalloc(&Y, NC * comp * 8) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(X, 0, NC * comp * 8);
  CUtensorMap in, out, inw;
  const cuuint64_t dims[4] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)(n + 2), NC};
  const cuuint64_t strides[3] = {(cuuint64_t)n * 8, (cuuint64_t)n * n * 8, (cuuint64_t)comp * 8};
  const cuuint32_t box[4] = {TX, TY, 1, 1}, es[4] = {1, 1, 1, 1};
  for (int w = 0; w < 2; ++w) {
    const CUresult r = ((encode_fn)fn)(w ? &out : &in, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, w ? (void*)Y : (void*)X, dims, strides, box, es,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
  }
  const int mode = argc > 2 ? atoi(argv[2]) : 0;  // one mode per process: a faulting kernel poisons the context
  const int ncl = argc > 3 ? atoi(argv[3]) : NC, shift = argc > 4 ? atoi(argv[4]) : 1;  // components moved per plane, pull shifts on/off
  {
    const cuuint32_t boxw[4] = {WX, TY, 1, 1};
    const CUresult r = ((encode_fn)fn)(&inw, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void*)X, dims, strides, boxw, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled (wide) failed: %d\n", (int)r); return 1; }
  }
  const double bytes = 16.0 * ncl * (n - 2 * EDGE_SKIP * TX) * (double)(n - 2 * EDGE_SKIP * TY) * n;
  double t;
  if (mode == 0) { t = run<0, 2>(in, inw, out, Y, n, lz, 5, ncl, shift); printf("n=%d TMA load + TMA store, %d comps, shift %d: %.3f ms  %.0f GB/s\n", n, ncl, shift, t, bytes / t / 1e6); }
  if (mode == 1) { t = run<1, 2>(in, inw, out, Y, n, lz, 5, ncl, shift); printf("n=%d TMA load + thread stores, %d comps, shift %d: %.3f ms  %.0f GB/s\n", n, ncl, shift, t, bytes / t / 1e6); }
  const cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
