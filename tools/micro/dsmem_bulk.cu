// Microbenchmark: how fast does cp.async.bulk move shared memory -> a peer CTA's shared memory inside a cluster of 4?
// Every CTA sends `bytes` to each of its 3 peers and receives 3 x `bytes`, like the per-step exchange of the LSTM kernels.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/dsmem_bulk tools/micro/dsmem_bulk.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) {
  uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o;
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c));
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t tx) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(tx) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_s2s(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(128, 1) k(int iters, uint32_t bytes, int pieces, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* src = smem;                       // 3 x bytes
  uint8_t* dst = smem + 3 * 16384 * 1;       // 3 x bytes (bytes <= 16384)
  __shared__ uint64_t bar;
  const uint32_t rank = ctarank();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = threadIdx.x; i < 3 * 16384; i += blockDim.x) src[i] = (uint8_t)(i + rank);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  cluster_sync();
  long long total = 0, worst = 0;
  for (int it = 0; it < iters; ++it) {
    if (threadIdx.x == 0) mbar_expect(&bar, 3 * bytes);
    cluster_sync();
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      for (uint32_t d = 1; d < 4; ++d) {
        const uint32_t peer = (rank + d) & 3;
        const uint32_t pb = bytes / pieces;
        for (int pc = 0; pc < pieces; ++pc)
          bulk_s2s(mapa(smem_u32(dst + (3 - d) * bytes + pc * pb), peer), smem_u32(src + (d - 1) * bytes + pc * pb), pb,
                   mapa(smem_u32(&bar), peer));
      }
      mbar_wait(&bar, it & 1);
      const long long dt = clock64() - t0;
      total += dt;
      if (dt > worst) worst = dt;
    }
  }
  cluster_sync();
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = total / iters; out[2 * blockIdx.x + 1] = worst; }
  // touch dst so the copies are not dead
  if (threadIdx.x == 1 && dst[5] == 77 && iters < 0) out[0] = 1;
}

int main() {
  long long* out;
  const int nblk = 128;
  cudaMalloc(&out, sizeof(long long) * 2 * nblk);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 16384 + 1024);
  for (int pieces : {1, 4}) {
    for (uint32_t bytes : {4096u, 8192u, 16384u}) {
      k<<<nblk, 128, 6 * 16384 + 1024>>>(200, bytes, pieces, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[2 * nblk];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mean = 0, mx = 0;
      for (int i = 0; i < nblk; ++i) { mean += h[2 * i]; if (h[2 * i + 1] > mx) mx = h[2 * i + 1]; }
      mean /= nblk;
      printf("bulk smem->peer smem: 3 x %u B out + 3 x %u B in per CTA, %d piece(s) each, %d CTAs: mean %lld cycles (%.1f B/clk/SM each way), worst %lld\n",
             bytes, bytes, pieces, nblk, mean, 3.0 * bytes / mean, mx);
    }
  }
  return 0;
}
