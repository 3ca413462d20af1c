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
  printf("%s\n", ok ? "PROBE ALL OK" : "PROBE HAD MISMATCHES");
  return ok ? 0 : 1;
}
