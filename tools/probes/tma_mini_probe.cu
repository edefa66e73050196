// Bisecting probe for tma_demand_probe.cu: which TMA form does this box accept?  One case per process (argv[1]).
//  0: mbarrier only   1: one 3-D f64 tile load   2: one 4-D f64 tile load   3: 4-D load + 4-D store   4: 3-D load of f32 (2x wide)
//  5: 4-D load with the descriptor copied to global memory
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap m3, const __grid_constant__ CUtensorMap m4, const __grid_constant__ CUtensorMap m3f,
                  const CUtensorMap* gm4, int mode, double* out) {
  __shared__ __align__(128) double tile[32 * 8];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = mode == 0 ? 0u : 2048u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
    if (mode == 1)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(s32(tile)), "l"(&m3), "r"(s32(&bar)), "r"(32), "r"(8), "r"(1) : "memory");
    if (mode == 2 || mode == 3)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(s32(tile)), "l"(&m4), "r"(s32(&bar)), "r"(32), "r"(8), "r"(1), "r"(1) : "memory");
    if (mode == 4)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(s32(tile)), "l"(&m3f), "r"(s32(&bar)), "r"(64), "r"(8), "r"(1) : "memory");
    if (mode == 5)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(s32(tile)), "l"(gm4), "r"(s32(&bar)), "r"(32), "r"(8), "r"(1), "r"(1) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(s32(&bar)) : "memory");
  if (mode == 3 && threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(&m4), "r"(s32(tile)), "r"(64), "r"(16), "r"(2), "r"(2) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  out[threadIdx.x] = tile[threadIdx.x];
}
typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  const int n = 128, nc = 4;
  const size_t comp = (size_t)(n + 2) * n * n;
  double *X, *out;
  cudaMalloc(&X, nc * comp * 8);
  cudaMalloc(&out, 256 * 8);
  std::vector<double> h(nc * comp);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (double)(i % 1000003);
  cudaMemcpy(X, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  CUtensorMap m3, m4, m3f;
  {
    const cuuint64_t d3[3] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)(n + 2) * nc}, s3[2] = {(cuuint64_t)n * 8, (cuuint64_t)n * n * 8};
    const cuuint32_t b3[3] = {32, 8, 1}, e3[3] = {1, 1, 1};
    printf("encode 3d f64: %d\n", (int)((encode_fn)fn)(&m3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, X, d3, s3, b3, e3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
    const cuuint64_t d4[4] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)(n + 2), nc}, s4[3] = {(cuuint64_t)n * 8, (cuuint64_t)n * n * 8, (cuuint64_t)comp * 8};
    const cuuint32_t b4[4] = {32, 8, 1, 1}, e4[4] = {1, 1, 1, 1};
    printf("encode 4d f64: %d\n", (int)((encode_fn)fn)(&m4, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, X, d4, s4, b4, e4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
    const cuuint64_t d3f[3] = {(cuuint64_t)2 * n, (cuuint64_t)n, (cuuint64_t)(n + 2) * nc};
    const cuuint32_t b3f[3] = {64, 8, 1};
    printf("encode 3d f32: %d\n", (int)((encode_fn)fn)(&m3f, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, X, d3f, s3, b3f, e3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
  }
  CUtensorMap* gm4;
  cudaMalloc(&gm4, sizeof(CUtensorMap));
  cudaMemcpy(gm4, &m4, sizeof m4, cudaMemcpyHostToDevice);
  k<<<1, 256>>>(m3, m4, m3f, gm4, mode, out);
  const cudaError_t e = cudaDeviceSynchronize();
  double r[4] = {0, 0, 0, 0};
  if (e == cudaSuccess) cudaMemcpy(r, out, sizeof r, cudaMemcpyDeviceToHost);
  // expected first element of the tile: X[c=1 (modes 2,3,5) or plane index 1][z=1][y=8][x=32]
  printf("mode %d: %s  tile[0..1] = %.0f %.0f\n", mode, cudaGetErrorString(e), r[0], r[1]);
  return 0;
}
