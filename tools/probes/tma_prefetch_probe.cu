// Probe: does cp.async.bulk.prefetch.tensor (L2 prefetch through a TMA descriptor) run on this box, and in which form?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_prefetch_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ void pf3(const CUtensorMap* map, int x, int y, int z) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(x), "r"(y), "r"(z) : "memory");
}
__global__ void k_param(const __grid_constant__ CUtensorMap m, int x, int y, int z, int lanes) {
  if ((int)threadIdx.x < lanes) pf3(&m, x + (lanes > 1 ? (int)threadIdx.x : 0), y, z);
}
__global__ void k_global(const CUtensorMap* m, int x, int y, int z) {
  if (threadIdx.x == 0) pf3(m, x, y, z);
}
typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void report(const char* what) {
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-60s %s\n", what, cudaGetErrorString(e));
  if (e != cudaSuccess) exit(0);
}
int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  printf("entry point %p query %d\n", fn, (int)q);
  const int nx = 512, ny = 512, nz = 64;
  double* X;
  cudaMalloc(&X, (size_t)nx * ny * nz * 8);
  for (int dtype = 0; dtype < 2; ++dtype) {
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)(dtype ? nx : 2 * nx), ny, nz};
    cuuint64_t strides[2] = {(cuuint64_t)nx * 8, (cuuint64_t)nx * ny * 8};
    cuuint32_t box[3] = {(cuuint32_t)(dtype ? 32 : 64), 8, 1}, es[3] = {1, 1, 1};
    CUresult r = ((encode_fn)fn)(&m, dtype ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, X, dims, strides, box, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode dtype=%s -> %d\n", dtype ? "f64" : "f32", (int)r);
    k_param<<<1, 32>>>(m, 0, 0, 0, 1);
    report("param-space descriptor, lane 0, in-bounds");
    k_param<<<1, 32>>>(m, -1, -1, 3, 1);
    report("param-space descriptor, lane 0, negative start");
    k_param<<<1, 32>>>(m, 0, 0, 0, 4);
    report("param-space descriptor, 4 lanes, different x");
    CUtensorMap* dm;
    cudaMalloc(&dm, sizeof m);
    cudaMemcpy(dm, &m, sizeof m, cudaMemcpyHostToDevice);
    k_global<<<1, 32>>>(dm, 0, 0, 0);
    report("global-memory descriptor, lane 0");
  }
  return 0;
}
