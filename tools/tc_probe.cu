// tc_probe: validates the tcgen05 / TMA / DSMEM encodings the lfmq kernels rely on, and measures the
// handful of on-chip rates the persistent-LSTM design depends on.  Built by `make tools` into
// tools/tc_probe, run on the GPU box:  tools/tc_probe [test ...]
//
//   gemm_k128   K-major, SWIZZLE_128B, operands written by threads   (layout of the h tile / U slice)
//   gemm_k128t  same, operands loaded by TMA
//   gemm_k64    K-major, SWIZZLE_64B, K-tile 32 (layout of the x tile / W slice), threads + TMA
//   gemm_mn     MN-major, SWIZZLE_128B via TMA (layout of the weight-gradient GEMM)
//   dsmem       cluster-of-4 h exchange: st.shared::cluster + remote mbarrier arrive, bytes/cycle
//   tmem_ld     tcgen05.ld 32x32b.x32 drain rate of a 128x256 fp32 accumulator
//   mufu        tanh.approx.f32 throughput
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../lfm_quant_b200/csrc/sm100.cuh"

using namespace lfmq::sm100;

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d: %s\n", cudaGetErrorString(e_), __FILE__, __LINE__, #x);  \
      exit(2);                                                                                 \
    }                                                                                          \
  } while (0)

// ------------------------------------------------------------------------------------------
// One-CTA GEMM  D[128 x N] = A * B^T  (fp32 accumulate in TMEM), everything resident in smem.
// ------------------------------------------------------------------------------------------
struct GemmParams {
  int N, K;             // K total
  int mode;             // 0: K-major SW128 manual, 1: K-major SW128 TMA, 2: K-major SW64 manual, 3: K-major SW64 TMA,
                        // 4: MN-major SW128 TMA
  const __nv_bfloat16* A;   // K-major: [128][K];  MN-major: [K][128]
  const __nv_bfloat16* B;   // K-major: [N][K];    MN-major: [K][N]
  float* D;                 // [128][N]
  uint32_t idesc;
  uint32_t lbo_a, sbo_a, lbo_b, sbo_b;   // bytes
};

__global__ void __launch_bounds__(128, 1) probe_gemm_kernel(GemmParams p, const __grid_constant__ CUtensorMap tma_a,
                                                            const __grid_constant__ CUtensorMap tma_b) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool kmajor = p.mode <= 3;
  const bool sw64 = (p.mode == 2 || p.mode == 3);
  const bool use_tma = (p.mode == 1 || p.mode == 3 || p.mode == 4);
  const int kblk = sw64 ? 32 : 64;                 // K elements per smem tile
  const int nkb = p.K / kblk;
  const int row_bytes = sw64 ? 64 : 128;
  const uint32_t a_tile = kmajor ? 128 * row_bytes : 2 * 64 * 128;          // bytes per k-block
  const uint32_t b_tile = kmajor ? p.N * row_bytes : (p.N / 64) * 64 * 128;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)nkb * a_tile;

  if (tid == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (use_tma) {
    if (tid == 0) {
      mbar_arrive_expect_tx(&bar_tma, (uint32_t)nkb * (a_tile + b_tile));
      for (int kb = 0; kb < nkb; ++kb) {
        if (kmajor) {
          tma_load_2d(sA + (size_t)kb * a_tile, &tma_a, &bar_tma, kb * kblk, 0);
          // B rows in chunks of <=256 rows per box
          tma_load_2d(sB + (size_t)kb * b_tile, &tma_b, &bar_tma, kb * kblk, 0);
        } else {
          for (int mb = 0; mb < 2; ++mb)
            tma_load_2d(sA + (size_t)kb * a_tile + mb * 8192, &tma_a, &bar_tma, mb * 64, kb * 64);
          for (int nb = 0; nb < p.N / 64; ++nb)
            tma_load_2d(sB + (size_t)kb * b_tile + nb * 8192, &tma_b, &bar_tma, nb * 64, kb * 64);
        }
      }
    }
    mbar_wait(&bar_tma, 0);
  } else {
    // thread-written swizzled K-major tiles: 16-byte chunk c of row r goes to r*row_bytes + ((c ^ f(r)) * 16)
    const int cpr = row_bytes / 16;   // chunks per row
    for (int kb = 0; kb < nkb; ++kb) {
      for (int idx = tid; idx < 128 * cpr; idx += 128) {
        const int r = idx / cpr, c = idx % cpr;
        const int sw = sw64 ? ((r >> 1) & 3) : (r & 7);
        const uint4 v = *reinterpret_cast<const uint4*>(p.A + (size_t)r * p.K + kb * kblk + c * 8);
        *reinterpret_cast<uint4*>(sA + (size_t)kb * a_tile + r * row_bytes + ((c ^ sw) << 4)) = v;
      }
      for (int idx = tid; idx < p.N * cpr; idx += 128) {
        const int r = idx / cpr, c = idx % cpr;
        const int sw = sw64 ? ((r >> 1) & 3) : (r & 7);
        const uint4 v = *reinterpret_cast<const uint4*>(p.B + (size_t)r * p.K + kb * kblk + c * 8);
        *reinterpret_cast<uint4*>(sB + (size_t)kb * b_tile + r * row_bytes + ((c ^ sw) << 4)) = v;
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
  }

  if (tid == 0) {
    tcgen05_fence_after();
    const uint32_t layout = sw64 ? LAYOUT_SW64 : LAYOUT_SW128;
    int first = 1;
    for (int kb = 0; kb < nkb; ++kb) {
      for (int k16 = 0; k16 < kblk / 16; ++k16) {
        const uint32_t koff = kmajor ? (uint32_t)k16 * 32 : (uint32_t)k16 * 2048;
        const uint64_t da = make_smem_desc(smem_u32(sA + (size_t)kb * a_tile) + koff, p.lbo_a, p.sbo_a, layout);
        const uint64_t db = make_smem_desc(smem_u32(sB + (size_t)kb * b_tile) + koff, p.lbo_b, p.sbo_b, layout);
        umma_f16(tmem, da, db, p.idesc, first ? 0u : 1u);
        first = 0;
      }
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tcgen05_fence_after();
  // epilogue: warp w reads TMEM lanes 32w..32w+31 (rows), 32 columns at a time
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < p.N; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.D[(size_t)row * p.N + c0 + j] = __uint_as_float(v[j]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

typedef CUresult (*PFN_cuTensorMapEncodeTiled_v12000)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                                      const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
}

static CUtensorMap make_map_2d(const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner, uint32_t box_outer,
                               CUtensorMapSwizzle sw) {
  static PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
  CUtensorMap m;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    exit(3);
  }
  return m;
}

static float bf16r(float v) { return __bfloat162float(__float2bfloat16(v)); }

static bool run_gemm(const char* name, int mode, int N, int K, uint32_t lbo_a, uint32_t sbo_a, uint32_t lbo_b,
                     uint32_t sbo_b) {
  const bool kmajor = mode <= 3;
  const bool sw64 = (mode == 2 || mode == 3);
  std::vector<float> A(128 * K), B(N * K);       // logical A[m][k], B[n][k]
  srand(1234 + mode);
  for (auto& v : A) v = bf16r((rand() % 2001 - 1000) / 1000.0f);
  for (auto& v : B) v = bf16r((rand() % 2001 - 1000) / 1000.0f);
  std::vector<__nv_bfloat16> hA(128 * K), hB(N * K);
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < K; ++k) hA[kmajor ? m * K + k : k * 128 + m] = __float2bfloat16(A[m * K + k]);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) hB[kmajor ? n * K + k : k * N + n] = __float2bfloat16(B[n * K + k]);
  __nv_bfloat16 *dA, *dB;
  float* dD;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dD, 128 * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, 128 * N * 4));
  GemmParams p;
  p.N = N; p.K = K; p.mode = mode; p.A = dA; p.B = dB; p.D = dD;
  p.idesc = make_idesc_bf16(128, N, !kmajor, !kmajor);
  p.lbo_a = lbo_a; p.sbo_a = sbo_a; p.lbo_b = lbo_b; p.sbo_b = sbo_b;
  CUtensorMap ma, mb;
  if (kmajor) {
    const int kb = sw64 ? 32 : 64;
    const CUtensorMapSwizzle sw = sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    ma = make_map_2d(dA, K, 128, kb, 128, sw);
    mb = make_map_2d(dB, K, N, kb, N, sw);
  } else {
    ma = make_map_2d(dA, 128, K, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    mb = make_map_2d(dB, N, K, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  const size_t smem = (size_t)(128 + N) * K * 2 + 2048;
  CK(cudaFuncSetAttribute(probe_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_gemm_kernel<<<1, 128, smem>>>(p, ma, mb);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[%s] KERNEL FAILED: %s\n", name, cudaGetErrorString(e));
    exit(4);   // context is dead
  }
  std::vector<float> D(128 * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
      const double d = fabs(s - D[m * N + n]);
      if (!(d <= maxerr)) maxerr = d;   // NaN-safe
      if (fabs(s) > maxref) maxref = fabs(s);
    }
  const bool ok = maxerr < 1e-3 * maxref;
  printf("[%s] mode=%d N=%d K=%d idesc=0x%08x lbo/sbo A=%u/%u B=%u/%u  max|err|=%.3e (max|ref|=%.2f)  %s\n", name,
         mode, N, K, p.idesc, lbo_a, sbo_a, lbo_b, sbo_b, maxerr, maxref, ok ? "OK" : "MISMATCH");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return ok;
}

// ------------------------------------------------------------------------------------------
// DSMEM exchange: cluster of 4, every CTA pushes a [128 x 64] bf16 slice (16 KB) to each peer
// with 16-byte st.shared::cluster stores, then arrives on the peers' mbarriers.
// ------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(128, 1) probe_dsmem_kernel(int iters, long long* cycles,
                                                                                         int* errors) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* buf = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bar, 4 * 128);     // every thread of every CTA (incl. self) arrives once per round
    fence_mbar_init();
  }
  cluster_sync_all();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // thread = row; writes its 128-byte row segment (8 chunks) into slot `rank` of every CTA's buffer
    for (uint32_t dst = 0; dst < 4; ++dst) {
      const uint32_t remote = mapa_u32(smem_u32(buf + rank * 16384 + tid * 128), dst);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t v = (uint32_t)(it * 131 + rank * 17 + tid * 3 + c);
        st_cluster_v4(remote + c * 16, v, v + 1, v + 2, v + 3);
      }
      mbar_arrive_cluster(mapa_u32(smem_u32(&bar), dst));   // release.cluster
    }
    mbar_wait_cluster(&bar, it & 1);                          // acquire.cluster
    // verify one word from each source
    if (it == iters - 1) {
      for (uint32_t src = 0; src < 4; ++src) {
        const uint32_t got = *reinterpret_cast<uint32_t*>(buf + src * 16384 + tid * 128 + 5 * 16);
        const uint32_t want = (uint32_t)(it * 131 + src * 17 + tid * 3 + 5);
        if (got != want) atomicAdd(errors, 1);
      }
    }
    cluster_sync_all();   // keep rounds from overlapping (buffer reuse); its cost is reported separately
  }
  long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  // cost of the bare cluster barrier
  long long t2 = clock64();
  for (int it = 0; it < iters; ++it) cluster_sync_all();
  long long t3 = clock64();
  if (tid == 0 && blockIdx.x == 0) cycles[1] = t3 - t2;
}

// tcgen05.ld drain rate + MUFU rate
__global__ void __launch_bounds__(128, 1) probe_tmem_ld_kernel(int iters, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s + ((uint32_t)(warp * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c0 = 0; c0 < 256; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += __uint_as_float(v[j] & 0x3fffffffu);
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
  sink[threadIdx.x] = acc;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 256);
}

__global__ void __launch_bounds__(128, 1) probe_mufu_kernel(int iters, long long* cycles, float* sink) {
  float a[8];
  for (int j = 0; j < 8; ++j) a[j] = 0.01f * (threadIdx.x + j);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = tanh_approx(a[j] + 0.25f);
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
  float s = 0;
  for (int j = 0; j < 8; ++j) s += a[j];
  sink[threadIdx.x] = s;
}

// ------------------------------------------------------------------------------------------
// v2 probes
// ------------------------------------------------------------------------------------------
// DSMEM exchange with the bulk-copy engine: every CTA stages its 16 KB slice locally, then ONE thread pushes it
// to the 3 peers with cp.async.bulk.shared::cluster; arrival is counted in bytes on the receivers' mbarriers.
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(128, 1) probe_dsmem_bulk_kernel(int iters,
                                                                                              long long* cycles,
                                                                                              int* errors) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* buf = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  cluster_sync_all();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint8_t* mine = buf + rank * 16384;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint32_t v = (uint32_t)(it * 131 + rank * 17 + tid * 3 + c);
      *reinterpret_cast<uint4*>(mine + tid * 128 + c * 16) = make_uint4(v, v + 1, v + 2, v + 3);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      mbar_arrive_expect_tx(&bar, 3 * 16384);
      for (uint32_t d = 1; d < 4; ++d) {
        const uint32_t dst = (rank + d) & 3;
        bulk_copy_s2c(mapa_u32(smem_u32(mine), dst), smem_u32(mine), 16384, mapa_u32(smem_u32(&bar), dst));
      }
    }
    mbar_wait_cluster(&bar, it & 1);
    if (it == iters - 1) {
      for (uint32_t src = 0; src < 4; ++src) {
        const uint32_t got = *reinterpret_cast<uint32_t*>(buf + src * 16384 + tid * 128 + 5 * 16);
        if (got != (uint32_t)(it * 131 + src * 17 + tid * 3 + 5)) atomicAdd(errors, 1);
      }
    }
    cluster_sync_all();
  }
  long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

// st.shared::cluster with lane-consecutive 16-byte chunks (fully coalesced remote stores)
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(128, 1) probe_dsmem_coal_kernel(int iters,
                                                                                              long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* buf = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bar, 4 * 128);
    fence_mbar_init();
  }
  cluster_sync_all();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (uint32_t dst = 0; dst < 4; ++dst) {
      const uint32_t remote = mapa_u32(smem_u32(buf + rank * 16384), dst);
#pragma unroll
      for (int c = 0; c < 8; ++c) st_cluster_v4(remote + (c * 128 + tid) * 16, it, tid, c, dst);
      mbar_arrive_cluster(mapa_u32(smem_u32(&bar), dst));
    }
    mbar_wait_cluster(&bar, it & 1);
    cluster_sync_all();
  }
  long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

// TMA multicast from global/L2: every CTA of the cluster fetches one 16 KB k-block of a [128 x 256] bf16 tile
// and multicasts it to all 4 CTAs (64 KB lands in each SM).
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(128, 1)
    probe_mcast_kernel(int iters, long long* cycles, const __grid_constant__ CUtensorMap tm, int* errors) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* buf = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x;
  const int tile = blockIdx.x / 4;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  cluster_sync_all();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (tid == 0) {
      mbar_arrive_expect_tx(&bar, 65536);
      tma_load_2d_mcast(buf + rank * 16384, &tm, &bar, rank * 64, tile * 128, 0xF);
    }
    mbar_wait_cluster(&bar, it & 1);
    cluster_sync_all();
  }
  long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  // element (row=tid, k = 64*kb + 0) of the tile is at swizzled chunk (0 ^ (tid&7))
  for (int kb = 0; kb < 4; ++kb) {
    const __nv_bfloat16 v = *reinterpret_cast<__nv_bfloat16*>(buf + kb * 16384 + tid * 128 + ((0 ^ (tid & 7)) << 4));
    const float want = (float)(((tile * 128 + tid) * 7 + (kb * 64) * 3) % 61);
    if (__bfloat162float(v) != want) atomicAdd(errors, 1);
  }
}

// tcgen05.ld variants
template <int MODE>   // 0: 4 warps, 4 x (x16) per wait ; 1: 8 warps (2 per lane quadrant), 4 x (x16) per wait
__global__ void __launch_bounds__(256, 1) probe_tmem_ld2_kernel(int iters, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const int nw = (MODE == 0) ? 4 : 8;
  const uint32_t tmem = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nw) {
    const int cbeg = (MODE == 0) ? 0 : (warp >> 2) * 128;
    const int cend = (MODE == 0) ? 256 : cbeg + 128;
    for (int it = 0; it < iters; ++it) {
      for (int c0 = cbeg; c0 < cend; c0 += 64) {
        uint32_t v[4][16];
#pragma unroll
        for (int q = 0; q < 4; ++q) tmem_ld_32x32b_x16(tmem + c0 + q * 16, v[q]);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int j = 0; j < 16; ++j) acc += __uint_as_float(v[q][j] & 0x3fffffffu);
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
  sink[threadIdx.x] = acc;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 256);
}

// MMA issue->completion for the per-step gate GEMM: NK x (M128 N256 K16), operands = whatever is in smem.
__global__ void __launch_bounds__(128, 1) probe_mma_lat_kernel(int nk, int reps, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (64 + 128) * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, 256, false, false);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int k = 0; k < nk; ++k) {
        const uint32_t kb = (k / 4) % 4, k16 = k % 4;
        const uint64_t da = make_smem_desc(smem_u32(smem + kb * 16384) + k16 * 32, 0, 1024, LAYOUT_SW128);
        const uint64_t db = make_smem_desc(smem_u32(smem + 65536 + kb * 32768) + k16 * 32, 0, 1024, LAYOUT_SW128);
        umma_f16(tmem_base_s + (r & 1) * 256, da, db, idesc, k > 0);
      }
      umma_commit(&bar);
      mbar_wait(&bar, r & 1);
    }
    long long t1 = clock64();
    cycles[0] = t1 - t0;
    // throughput: issue everything, wait once
    t0 = clock64();
    for (int r = 0; r < reps; ++r)
      for (int k = 0; k < nk; ++k) {
        const uint32_t kb = (k / 4) % 4, k16 = k % 4;
        const uint64_t da = make_smem_desc(smem_u32(smem + kb * 16384) + k16 * 32, 0, 1024, LAYOUT_SW128);
        const uint64_t db = make_smem_desc(smem_u32(smem + 65536 + kb * 32768) + k16 * 32, 0, 1024, LAYOUT_SW128);
        umma_f16(tmem_base_s + (r & 1) * 256, da, db, idesc, k > 0);
      }
    umma_commit(&bar);
    mbar_wait(&bar, reps & 1);
    t1 = clock64();
    cycles[1] = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

__global__ void __launch_bounds__(128, 1) probe_mufu2_kernel(int iters, long long* cycles, float* sink) {
  uint32_t a[8];
  for (int j = 0; j < 8; ++j) a[j] = pack_bf16x2(0.01f * (threadIdx.x + j), 0.02f * j);
  const uint32_t q = pack_bf16x2(0.25f, 0.25f), one = pack_bf16x2(1.f, 1.f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = tanh_bf16x2(fma_bf16x2(a[j], one, q));
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
  float s = 0;
  for (int j = 0; j < 8; ++j) s += bf16_lo(a[j]) + bf16_hi(a[j]);
  sink[threadIdx.x] = s;
}

// ---- 3-D TMA store of a K-major SW128 tile (128-byte rows = 4 gates x 16 units) into dz laid out as
// [b][t][16-unit block][gate][16]: the row of a chunk is 128 contiguous bytes.  (A 5-D map over the gate-major layout
// -- box 16 x 1 x 4 x 1 x 128, SWIZZLE_128B -- raised an illegal memory access on B200 and is not used.) ----
__global__ void __launch_bounds__(128, 1) probe_st3d_kernel(const __grid_constant__ CUtensorMap tm, int ublk, int t,
                                                            int b0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int m = threadIdx.x;
  for (int c = 0; c < 8; ++c) {          // 16-byte chunk c = 2*gate + half, XOR-swizzled with the row like the A operand
    uint16_t v[8];
    for (int e = 0; e < 8; ++e) v[e] = (uint16_t)(m * 64 + (c / 2) * 16 + (c % 2) * 8 + e);
    *reinterpret_cast<uint4*>(smem + m * 128 + ((c ^ (m & 7)) << 4)) = *reinterpret_cast<uint4*>(v);
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (m == 0) {
    tma_store_3d(&tm, smem, ublk * 64, t, b0);
    bulk_commit_group();
    bulk_wait_group0();
  }
}

static bool run_st3d() {
  const int Bcap = 300, B = 150, T1 = 3;
  const size_t n = (size_t)Bcap * T1 * 1024;
  uint16_t* d;
  CK(cudaMalloc(&d, n * 2));
  CK(cudaMemset(d, 0xFF, n * 2));
  static PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
  CUtensorMap tm;
  cuuint64_t dims[3] = {1024, (cuuint64_t)T1, (cuuint64_t)B};
  cuuint64_t strides[2] = {2048, (cuuint64_t)2048 * T1};
  cuuint32_t box[3] = {64, 1, 128};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("[st3d] encode failed %d\n", (int)r); return false; }
  CK(cudaFuncSetAttribute(probe_st3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 1024));
  probe_st3d_kernel<<<1, 128, 16384 + 1024>>>(tm, 5, 1, 0);
  probe_st3d_kernel<<<1, 128, 16384 + 1024>>>(tm, 9, 2, 128);
  CK(cudaDeviceSynchronize());
  std::vector<uint16_t> h(n);
  CK(cudaMemcpy(h.data(), d, n * 2, cudaMemcpyDeviceToHost));
  long bad = 0, written = 0, first = -1;
  for (int b = 0; b < Bcap; ++b)
    for (int t = 0; t < T1; ++t)
      for (int col = 0; col < 1024; ++col) {
        const int ub = col / 64, g = (col % 64) / 16, jj = col % 16;
        uint16_t want = 0xFFFF;
        if (t == 1 && ub == 5 && b < 128) want = (uint16_t)(b * 64 + g * 16 + jj);
        if (t == 2 && ub == 9 && b >= 128 && b < B) want = (uint16_t)((b - 128) * 64 + g * 16 + jj);
        const uint16_t got = h[((size_t)b * T1 + t) * 1024 + col];
        if (got != 0xFFFF) ++written;
        if (got != want) { if (first < 0) first = ((long)b * T1 + t) * 1024 + col; ++bad; }
      }
  printf("[st3d] elements written %ld (expected %d), mismatches %ld", written, 128 * 64 + 22 * 64, bad);
  if (first >= 0) {
    const long e = first;
    printf(", first at b=%ld t=%ld col=%ld got %u", e / (T1 * 1024), (e / 1024) % T1, e % 1024, (unsigned)h[e]);
  }
  printf("\n");
  CK(cudaFree(d));
  return bad == 0;
}

static bool want(int argc, char** argv, const char* name) {
  if (argc <= 1) return true;
  for (int i = 1; i < argc; ++i)
    if (!strcmp(argv[i], name)) return true;
  return false;
}

int main(int argc, char** argv) {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs, smem/block optin %zu\n", prop.name, prop.major, prop.minor,
         prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
  bool ok = true;
  if (want(argc, argv, "gemm_k128")) {
    ok &= run_gemm("gemm_k128 N=256 K=64", 0, 256, 64, 0, 1024, 0, 1024);
    ok &= run_gemm("gemm_k128 N=256 K=256", 0, 256, 256, 0, 1024, 0, 1024);
    ok &= run_gemm("gemm_k128 N=64 K=128", 0, 64, 128, 0, 1024, 0, 1024);
  }
  if (want(argc, argv, "gemm_k128t")) ok &= run_gemm("gemm_k128t N=256 K=256", 1, 256, 256, 0, 1024, 0, 1024);
  if (want(argc, argv, "gemm_k64")) {
    ok &= run_gemm("gemm_k64 manual N=256 K=32", 2, 256, 32, 0, 512, 0, 512);
    ok &= run_gemm("gemm_k64 tma    N=256 K=64", 3, 256, 64, 0, 512, 0, 512);
  }
  if (want(argc, argv, "gemm_mn")) {
    ok &= run_gemm("gemm_mn N=256 K=64", 4, 256, 64, 8192, 1024, 8192, 1024);
    ok &= run_gemm("gemm_mn N=256 K=128", 4, 256, 128, 8192, 1024, 8192, 1024);
  }
  long long* dcy;
  float* dsink;
  int* derr;
  CK(cudaMalloc(&dcy, 64));
  CK(cudaMalloc(&dsink, 4096));
  CK(cudaMalloc(&derr, 4));
  long long cy[2];
  if (want(argc, argv, "dsmem")) {
    const int iters = 200;
    CK(cudaMemset(derr, 0, 4));
    CK(cudaFuncSetAttribute(probe_dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384 + 2048));
    probe_dsmem_kernel<<<4 * 32, 128, 4 * 16384 + 2048>>>(iters, dcy, derr);
    CK(cudaDeviceSynchronize());
    int herr;
    CK(cudaMemcpy(cy, dcy, 16, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
    const double per = (double)cy[0] / iters, sync = (double)cy[1] / iters;
    printf("[dsmem] 32 clusters x 4 CTAs: round (4 x 16 KB pushes + mbarrier + cluster.sync) = %.0f cyc, bare "
           "cluster.sync = %.0f cyc -> exchange ~%.0f cyc, %.1f B/cyc/SM outgoing (48 KB remote), errors=%d\n",
           per, sync, per - sync, 49152.0 / (per - sync), herr);
    ok &= (herr == 0);
  }
  if (want(argc, argv, "tmem_ld")) {
    const int iters = 200;
    probe_tmem_ld_kernel<<<148, 128>>>(iters, dcy, dsink);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(cy, dcy, 8, cudaMemcpyDeviceToHost));
    printf("[tmem_ld] 128x256 fp32 drain (+256 FADD/thread): %.0f cyc per tile -> %.1f B/cyc/SM\n",
           (double)cy[0] / iters, 131072.0 / ((double)cy[0] / iters));
  }
  if (want(argc, argv, "mufu")) {
    const int iters = 1000;
    probe_mufu_kernel<<<148, 128>>>(iters, dcy, dsink);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(cy, dcy, 8, cudaMemcpyDeviceToHost));
    printf("[mufu] tanh.approx: %.2f per cycle per SM (128 threads, 8-way ILP)\n",
           128.0 * 8 * iters / (double)cy[0]);
  }
  if (want(argc, argv, "dsmem_bulk")) {
    const int iters = 200;
    CK(cudaMemset(derr, 0, 4));
    CK(cudaFuncSetAttribute(probe_dsmem_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384 + 2048));
    probe_dsmem_bulk_kernel<<<4 * 32, 128, 4 * 16384 + 2048>>>(iters, dcy, derr);
    CK(cudaDeviceSynchronize());
    int herr;
    CK(cudaMemcpy(cy, dcy, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
    const double per = (double)cy[0] / iters;
    printf("[dsmem_bulk] round (local 16 KB stage + 3 x 16 KB cp.async.bulk s2c + wait + cluster.sync) = %.0f cyc "
           "(minus ~406 sync) -> %.1f B/cyc/SM outgoing, errors=%d\n", per, 49152.0 / (per - 406), herr);
    ok &= (herr == 0);
  }
  if (want(argc, argv, "dsmem_coal")) {
    const int iters = 200;
    CK(cudaFuncSetAttribute(probe_dsmem_coal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384 + 2048));
    probe_dsmem_coal_kernel<<<4 * 32, 128, 4 * 16384 + 2048>>>(iters, dcy);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(cy, dcy, 8, cudaMemcpyDeviceToHost));
    const double per = (double)cy[0] / iters;
    printf("[dsmem_coal] lane-consecutive st.shared::cluster.v4: round = %.0f cyc -> %.1f B/cyc/SM outgoing\n", per,
           49152.0 / (per - 406));
  }
  if (want(argc, argv, "mcast")) {
    const int iters = 200, tiles = 32;
    std::vector<__nv_bfloat16> hh((size_t)tiles * 128 * 256);
    for (int r = 0; r < tiles * 128; ++r)
      for (int k = 0; k < 256; ++k) hh[(size_t)r * 256 + k] = __float2bfloat16((float)((r * 7 + k * 3) % 61));
    __nv_bfloat16* dh;
    CK(cudaMalloc(&dh, hh.size() * 2));
    CK(cudaMemcpy(dh, hh.data(), hh.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap tm = make_map_2d(dh, 256, tiles * 128, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    CK(cudaMemset(derr, 0, 4));
    CK(cudaFuncSetAttribute(probe_mcast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384 + 2048));
    probe_mcast_kernel<<<4 * tiles, 128, 4 * 16384 + 2048>>>(iters, dcy, tm, derr);
    CK(cudaDeviceSynchronize());
    int herr;
    CK(cudaMemcpy(cy, dcy, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
    printf("[mcast] 4 x 16 KB TMA multicast (L2 -> 4 CTAs, 64 KB per SM): round = %.0f cyc (incl ~406 sync), errors=%d\n",
           (double)cy[0] / iters, herr);
    ok &= (herr == 0);
    cudaFree(dh);
  }
  if (want(argc, argv, "tmem_ld2")) {
    const int iters = 200;
    probe_tmem_ld2_kernel<0><<<148, 256>>>(iters, dcy, dsink);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(cy, dcy, 8, cudaMemcpyDeviceToHost));
    printf("[tmem_ld2] 4 warps, 4 x x16 per wait: %.0f cyc per 128x256 tile\n", (double)cy[0] / iters);
    probe_tmem_ld2_kernel<1><<<148, 256>>>(iters, dcy, dsink);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(cy, dcy, 8, cudaMemcpyDeviceToHost));
    printf("[tmem_ld2] 8 warps, 4 x x16 per wait: %.0f cyc per 128x256 tile\n", (double)cy[0] / iters);
  }
  if (want(argc, argv, "mma_lat")) {
    const int reps = 20;
    CK(cudaFuncSetAttribute(probe_mma_lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int nk : {2, 16, 18}) {
      probe_mma_lat_kernel<<<148, 128, 200 * 1024>>>(nk, reps, dcy);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(cy, dcy, 16, cudaMemcpyDeviceToHost));
      printf("[mma_lat] %2d x (M128 N256 K16): issue->commit->wait %.0f cyc per batch; back-to-back %.0f cyc per batch "
             "(%.1f cyc/MMA)\n", nk, (double)cy[0] / reps, (double)cy[1] / reps, (double)cy[1] / reps / nk);
    }
  }
  if (want(argc, argv, "mufu2")) {
    const int iters = 1000;
    probe_mufu2_kernel<<<148, 128>>>(iters, dcy, dsink);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(cy, dcy, 8, cudaMemcpyDeviceToHost));
    printf("[mufu2] tanh.approx.bf16x2 (+1 HFMA2): %.2f instr = %.2f tanh per cycle per SM\n",
           128.0 * 8 * iters / (double)cy[0], 2 * 128.0 * 8 * iters / (double)cy[0]);
  }
  if (want(argc, argv, "st3d")) ok &= run_st3d();
  printf("%s\n", ok ? "PROBE ALL OK" : "PROBE HAD MISMATCHES");
  return ok ? 0 : 1;
}
