// Stand-alone check of the TMA tensor view das_tile uses: rows of 8-byte elements seen as (65, n_seg, rows) with a
// 512-byte segment stride (segments overlap by one element) so that a box may start at element 1.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void k(const __grid_constant__ CUtensorMap map, u64* out, int n_seg, int rows, int c0) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) u64 bar;
  unsigned dst = (unsigned)__cvta_generic_to_shared(sm), b = (unsigned)__cvta_generic_to_shared(&bar);
  int bytes = 64 * 8 * n_seg * rows;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(&map), "r"(c0), "r"(0), "r"(0), "r"(b) : "memory");
  }
  asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0; @p bra D; bra W; D: }" ::"r"(b) : "memory");
  for (int i = threadIdx.x; i < bytes / 8; i += blockDim.x) out[i] = ((u64*)sm)[i];
}

int main(int argc, char** argv) {
  int dim0 = argc > 1 ? atoi(argv[1]) : 65, c0 = argc > 2 ? atoi(argv[2]) : 1;
  const int n_seg = 7, rows = 8, row_bytes = 3584, n_rows = 16;
  std::vector<u64> h(n_rows * row_bytes / 8 + 1024);
  for (size_t i = 0; i < h.size(); i++) h[i] = i;
  u64 *d, *o; CK(cudaMalloc(&d, h.size() * 8)); CK(cudaMalloc(&o, 64 * n_seg * rows * 8));
  CK(cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn; cudaDriverEntryPointQueryResult q; CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  CUtensorMap map;
  cuuint64_t dims[3] = {(cuuint64_t)dim0, n_seg, n_rows}; cuuint64_t strides[2] = {512, row_bytes};
  cuuint32_t box[3] = {64, n_seg, rows}; cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("dim0 %d c0 %d encode rc %d\n", dim0, c0, (int)r);
  if (r != CUDA_SUCCESS) return 0;
  int smem = 64 * 8 * n_seg * rows;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k<<<1, 128, smem>>>(map, o, n_seg, rows, c0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 0;
  std::vector<u64> res(64 * n_seg * rows); CK(cudaMemcpy(res.data(), o, res.size() * 8, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int rr = 0; rr < rows; rr++) for (int i = 0; i < 64 * n_seg; i++) {
    u64 want = (u64)rr * row_bytes / 8 + i + c0;
    if (res[rr * 64 * n_seg + i] != want) { if (bad < 5) printf("row %d elem %d got %llu want %llu\n", rr, i, res[rr * 64 * n_seg + i], want); bad++; }
  }
  printf("mismatches %d of %zu\n", bad, res.size());
  return 0;
}
