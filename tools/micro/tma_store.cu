// Microbenchmark: cycles until the shared-memory read of a TMA store of a [128 rows x 128 B] box completes
// (cp.async.bulk.wait_group.read 0) and until the store itself completes (wait_group 0), as a function of the row
// stride in global memory; plus a 16 KB 1-D bulk store.  128 CTAs store concurrently, like the backward recurrence.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/tma_store tools/micro/tma_store.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tm, uint8_t* flat, int iters, int mode,
                                             long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) smem[i] = (uint8_t)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x != 0) return;
  long long t_read = 0, t_done = 0;
  for (int it = 0; it < iters; ++it) {
    const long long t0 = clock64();
    if (mode == 0) {          // tensor store: box 64 bf16 x 128 rows at (col 0, row blockIdx.x * 128 + ...)
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                       reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(smem)), "r"((it & 7) * 64), "r"((int)blockIdx.x * 128)
                   : "memory");
    } else {                  // 1-D bulk store of 16 KB
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                       reinterpret_cast<uint64_t>(flat + ((size_t)blockIdx.x * 8 + (it & 7)) * 16384)),
                   "r"(smem_u32(smem)), "r"(16384) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    const long long t1 = clock64();
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    const long long t2 = clock64();
    t_read += t1 - t0;
    t_done += t2 - t0;
  }
  out[2 * blockIdx.x] = t_read / iters;
  out[2 * blockIdx.x + 1] = t_done / iters;
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(fn);
  const int nblk = 128, rows = nblk * 128;
  long long* out;
  cudaMalloc(&out, sizeof(long long) * 2 * nblk);
  uint8_t* buf;
  const size_t max_stride = 49 * 1024 * 2;            // the kernels' dz: [B][T+1][1024] bf16 -> 100 352 B per row
  cudaMalloc(&buf, (size_t)rows * max_stride);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 1024);
  const size_t strides[] = {1024, 2048, 8192, max_stride};     // bytes between consecutive rows (>= 8 boxes of 128 B)
  for (size_t st : strides) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {512, (cuuint64_t)rows};
    cuuint64_t sb[1] = {st};
    cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, sb, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    k<<<nblk, 128, 16384 + 1024>>>(tm, buf, 200, 0, out);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
    long long h[2 * nblk];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    long long a = 0, b = 0;
    for (int i = 0; i < nblk; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
    printf("tensor store 128 rows x 128 B, row stride %7zu B: smem read done after %5lld cycles, store complete after %5lld (mean of %d CTAs)\n",
           st, a / nblk, b / nblk, nblk);
  }
  {
    CUtensorMap tm{};
    k<<<nblk, 128, 16384 + 1024>>>(tm, buf, 200, 1, out);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
    long long h[2 * nblk];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    long long a = 0, b = 0;
    for (int i = 0; i < nblk; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
    printf("1-D bulk store of 16 KB (contiguous):            smem read done after %5lld cycles, store complete after %5lld\n",
           a / nblk, b / nblk);
  }
  return 0;
}
