// Microbenchmark: per-SM throughput of the transcendental (MUFU/XU) pipe for the activation forms the
// tensor-core LSTM epilogue can use: tanh.approx.f32, tanh.approx.f16x2, tanh.approx.bf16x2, ex2.approx.f16x2,
// and an FMA-pipe polynomial tanh.  One CTA of 512 threads per SM, 8 independent chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_mufu ubench_mufu.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float tanh_f32(float x) { float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t tanh_h2(uint32_t x) { uint32_t y; asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t tanh_b2(uint32_t x) { uint32_t y; asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
// odd polynomial on a clamped argument (FMA/ALU pipes only) -- cost model, not tuned coefficients
__device__ __forceinline__ float tanh_poly(float x) {
  const float xc = fminf(fmaxf(x, -4.0f), 4.0f);
  const float s = xc * xc;
  float p = fmaf(s, -2.0e-6f, 1.1e-4f);
  p = fmaf(s, p, -2.4e-3f);
  p = fmaf(s, p, 2.6e-2f);
  p = fmaf(s, p, -1.6e-1f);
  p = fmaf(s, p, 0.5f);
  p = fmaf(s, p, 1.0f);
  return xc * p * 0.3f;
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) bench(float* out, int iters, long long* cyc) {
  float f[8];
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = 0.001f * (threadIdx.x + i); u[i] = 0x2c002e00u + threadIdx.x + i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) f[i] = tanh_f32(f[i]);
      if (MODE == 1) u[i] = tanh_h2(u[i]);
      if (MODE == 2) u[i] = tanh_b2(u[i]);
      if (MODE == 3) u[i] = ex2_h2(u[i]);
      if (MODE == 4) f[i] = tanh_poly(f[i]);
      if (MODE == 5) { f[i] = tanh_f32(f[i]); u[i] = tanh_h2(u[i]); }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += f[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 8);
  const char* names[] = {"tanh.approx.f32", "tanh.approx.f16x2", "tanh.approx.bf16x2", "ex2.approx.f16x2", "poly tanh (FMA pipe)", "f32 + f16x2 interleaved"};
  const int iters = 2000;
  for (int m = 0; m < 6; ++m) {
    for (int rep = 0; rep < 2; ++rep) {
      switch (m) {
        case 0: bench<0><<<148, 512>>>(out, iters, cyc); break;
        case 1: bench<1><<<148, 512>>>(out, iters, cyc); break;
        case 2: bench<2><<<148, 512>>>(out, iters, cyc); break;
        case 3: bench<3><<<148, 512>>>(out, iters, cyc); break;
        case 4: bench<4><<<148, 512>>>(out, iters, cyc); break;
        case 5: bench<5><<<148, 512>>>(out, iters, cyc); break;
      }
    }
    long long h = 0;
    cudaError_t e = cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("%s: %s\n", names[m], cudaGetErrorString(e)); return 1; }
    const double instr = (double)iters * 8 * 512 * (m == 5 ? 2 : 1);   // thread-level instructions per SM
    printf("%-28s %9lld cyc  -> %.2f thread-instr/clk/SM  (%.2f results/clk/SM)\n", names[m], h, instr / h,
           instr / h * ((m >= 1 && m <= 3) ? 2.0 : (m == 5 ? 1.5 : 1.0)));
  }
  return 0;
}
