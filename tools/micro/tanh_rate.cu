// Microbenchmark: throughput of the gate nonlinearity variants per SM (8 warps resident, like the forward epilogue).
//   f32   : tanh.approx.f32 per value
//   f16x2 : cvt.rn.f16x2.f32 (pack two) -> tanh.approx.f16x2 -> two cvt.f32.f16
//   bf16x2: cvt.rn.bf16x2.f32 -> tanh.approx.bf16x2 -> shift / mask unpack
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/tanh_rate tools/micro/tanh_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

__device__ __forceinline__ float tanh_f32(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void tanh_f16x2(float a, float b, float& ya, float& yb) {
  uint32_t p, q;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(b), "f"(a));
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(q) : "r"(p));
  __half2 h = *reinterpret_cast<__half2*>(&q);
  ya = __low2float(h);
  yb = __high2float(h);
}
__device__ __forceinline__ void tanh_bf16x2(float a, float b, float& ya, float& yb) {
  uint32_t p, q;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(b), "f"(a));
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(q) : "r"(p));
  ya = __uint_as_float(q << 16);
  yb = __uint_as_float(q & 0xffff0000u);
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(int iters, float* out, long long* cyc) {
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = 0.01f * (threadIdx.x + j) - 1.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      float a, b;
      if (MODE == 0) { a = tanh_f32(v[j]); b = tanh_f32(v[j + 1]); }
      if (MODE == 1) tanh_f16x2(v[j], v[j + 1], a, b);
      if (MODE == 2) tanh_bf16x2(v[j], v[j + 1], a, b);
      v[j] = fmaf(a, 0.9f, 0.05f);          // keep a dependent fp32 op per value, like the gate arithmetic
      v[j + 1] = fmaf(b, 0.9f, -0.05f);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  const char* names[3] = {"tanh.approx.f32", "f16x2 (pack + tanh.approx.f16x2 + 2 unpack)", "bf16x2 (pack + tanh.approx.bf16x2 + shift/mask)"};
  for (int mode = 0; mode < 3; ++mode) {
    if (mode == 0) k<0><<<148, 256>>>(iters, out, cyc);
    if (mode == 1) k<1><<<148, 256>>>(iters, out, cyc);
    if (mode == 2) k<2><<<148, 256>>>(iters, out, cyc);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0;
    for (int i = 0; i < 148; ++i) m += h[i];
    m /= 148;
    const double vals = (double)iters * 16 * 256;
    printf("%-52s %.1f values/clk/SM (8 warps, %.0f cycles)\n", names[mode], vals / m, m);
  }
  // accuracy of the three on a grid
  return 0;
}
