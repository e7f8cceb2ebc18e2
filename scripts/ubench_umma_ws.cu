// Microbenchmark: does tcgen05.mma.ws (weight-stationary: the B operand parked in a collector buffer, `collector::bN::fill`
// on the first MMA of a group, `::use` / `::lastuse` on the rest) remove the shared-memory re-reads of a B slab that several
// consecutive MMAs share?  In the recurrent kernel B is the activation slab (N sequences x 16 K) and A the streamed weights: the
// 8 row tiles of the gate contraction all multiply the SAME activation slab per K step.
// Pattern: groups of G MMAs (M=128, N=64, K=16) with G different A slabs and ONE B slab, accumulating into G TMEM tiles.
// Reports cycles per MMA for the plain instruction and for .ws, and checks that both leave identical accumulators (same TMEM layout).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_umma_ws ubench_umma_ws.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
// One asm block per group of G MMAs: one elect, the B descriptor built once, the A descriptor / accumulator column advanced by
// immediates (A slabs 256 B apart = 16 descriptor units, accumulator tiles 64 columns apart) -- the issue cost must stay below
// the ~32-48 cycles an MMA takes, or the benchmark measures the issuing thread instead of the tensor pipe.
#define MMA_T(INSTR, I)                                                                                         \
  "add.u32 dd, %0, " #I " * 64;\n\tadd.u32 aa, %1, " #I " * 16;\n\tmov.b64 da, {aa, %2};\n\t@q " INSTR " [dd], da, db, %5, p;\n\t"
#define PLAIN "tcgen05.mma.cta_group::1.kind::f16"
#define WS(OP) "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::" #OP
#define GROUP_ASM(BODY)                                                                                                                  \
  asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t.reg .b32 rx, dd, aa;\n\telect.sync rx|q, 0xffffffff;\n\tmov.b64 db, {%3, %4};\n\t" \
               "setp.ne.b32 p, %6, 0;\n\t" BODY "}\n" ::"r"(d),                                                                            \
               "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc)                                                              \
               : "memory")
template <int G, int V>
__device__ __forceinline__ void group(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
  if constexpr (V == 0) {
    if constexpr (G == 1) GROUP_ASM(MMA_T(PLAIN, 0));
    if constexpr (G == 2) GROUP_ASM(MMA_T(PLAIN, 0) MMA_T(PLAIN, 1));
    if constexpr (G == 4) GROUP_ASM(MMA_T(PLAIN, 0) MMA_T(PLAIN, 1) MMA_T(PLAIN, 2) MMA_T(PLAIN, 3));
    if constexpr (G == 8) GROUP_ASM(MMA_T(PLAIN, 0) MMA_T(PLAIN, 1) MMA_T(PLAIN, 2) MMA_T(PLAIN, 3) MMA_T(PLAIN, 4) MMA_T(PLAIN, 5) MMA_T(PLAIN, 6) MMA_T(PLAIN, 7));
  } else if constexpr (V == 1) {
    if constexpr (G == 1) GROUP_ASM(MMA_T(WS(discard), 0));
    if constexpr (G == 2) GROUP_ASM(MMA_T(WS(discard), 0) MMA_T(WS(discard), 1));
    if constexpr (G == 4) GROUP_ASM(MMA_T(WS(discard), 0) MMA_T(WS(discard), 1) MMA_T(WS(discard), 2) MMA_T(WS(discard), 3));
    if constexpr (G == 8) GROUP_ASM(MMA_T(WS(discard), 0) MMA_T(WS(discard), 1) MMA_T(WS(discard), 2) MMA_T(WS(discard), 3) MMA_T(WS(discard), 4) MMA_T(WS(discard), 5) MMA_T(WS(discard), 6) MMA_T(WS(discard), 7));
  } else {
    if constexpr (G == 1) GROUP_ASM(MMA_T(WS(discard), 0));
    if constexpr (G == 2) GROUP_ASM(MMA_T(WS(fill), 0) MMA_T(WS(lastuse), 1));
    if constexpr (G == 4) GROUP_ASM(MMA_T(WS(fill), 0) MMA_T(WS(use), 1) MMA_T(WS(use), 2) MMA_T(WS(lastuse), 3));
    if constexpr (G == 8) GROUP_ASM(MMA_T(WS(fill), 0) MMA_T(WS(use), 1) MMA_T(WS(use), 2) MMA_T(WS(use), 3) MMA_T(WS(use), 4) MMA_T(WS(use), 5) MMA_T(WS(use), 6) MMA_T(WS(lastuse), 7));
  }
}
__device__ __forceinline__ uint32_t dlo(uint32_t a, uint32_t lbo) { return ((a >> 4) & 0x3FFF) | ((lbo >> 4) << 16); }
__device__ __forceinline__ uint32_t dhi(uint32_t sbo) { return ((sbo >> 4) & 0x3FFF) | (1u << 14); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// variant: 0 plain, 1 ws with fill on every MMA (no reuse), 2 ws fill / use / lastuse per group
// A: two K-major no-swizzle 128 x 256 matrices (LBO 128, SBO 4096, K step 256 B) = 32 slabs; B: K-major no-swizzle 64 x 16 slabs of 2 KB
template <int G, int V>
__global__ void __launch_bounds__(128, 1) bench(int bmn, int n_groups, const __half* __restrict__ init, long long* out, float* acc_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < (128 + 16) * 1024 / 2; i += blockDim.x) reinterpret_cast<__half*>(smem)[i] = init[i];
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (threadIdx.x < 32) {
    const uint32_t N = 64, M = 128;
    // f16 x f16 -> f32; A K-major; B K-major (bmn = 0) or MN-major (bmn = 1: the recurrent kernel's activation tiles,
    // elem(k,n) at (k/8)*(N*16) + (n/8)*128 + (k%8)*16 + (n%8)*2, LBO = N*16, SBO = 128)
    const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)bmn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 128 * 1024;
    const uint32_t ahi = dhi(4096), bhi = bmn ? dhi(128) : dhi(256);   // K-major B slab: 8 n-groups of (2 k-cores x 128 B) = 256 B apart
    const uint32_t b_lbo = bmn ? N * 16 : 128;
    uint32_t parity = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
#pragma unroll 1
      for (int g = 0; g < n_groups; ++g) {
        const uint32_t blo = dlo(b0 + (uint32_t)(g & 7) * 2048u, b_lbo);
        const uint32_t alo = dlo(a0 + (uint32_t)(g & 1) * 65536u + (uint32_t)((g * G) & 15 & ~(G - 1)) * 256u, 128);   // G consecutive 4 KB slabs
        group<G, V>(tmem, alo, ahi, blo, bhi, idesc, g > 0);
      }
      const long long t1 = clock64();
      if (threadIdx.x == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      { long long tw = clock64(); while (!try_wait(smem_u32(&bar), parity)) { if (clock64() - tw > 200000000LL) __trap(); } }
      parity ^= 1;
      const long long t2 = clock64();
      if (threadIdx.x == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // accumulators of the last repetition: tile i = columns [64 i, 64 i + 64), lane = row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = 0; i < G; ++i)
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(i * 64 + c0), r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int c = 0; c < 16; ++c) acc_out[((size_t)i * 128 + warp * 32 + lane) * 64 + c0 + c] = __uint_as_float(r[c]);
    }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  const int n_half = (128 + 16) * 1024 / 2;
  __half* h_init = (__half*)malloc(n_half * sizeof(__half));
  srand(1);
  for (int i = 0; i < n_half; ++i) h_init[i] = __float2half((float)(rand() % 7 - 3) * 0.25f);   // exact in f16, sums exact in f32
  __half* d_init;
  long long* d_t;
  float* d_acc;
  cudaMalloc(&d_init, n_half * sizeof(__half));
  cudaMemcpy(d_init, h_init, n_half * sizeof(__half), cudaMemcpyHostToDevice);
  cudaMalloc(&d_t, 64);
  const size_t acc_n = (size_t)8 * 128 * 64;
  cudaMalloc(&d_acc, acc_n * sizeof(float));
  float* ref = (float*)malloc(acc_n * sizeof(float));
  float* got = (float*)malloc(acc_n * sizeof(float));
  const char* names[3] = {"plain", "ws, no reuse (discard)", "ws, fill/use/lastuse"};
  int bmn = 0;
  auto run = [&](int G, int v, auto kern) {
    const int n_groups = 512 / G;   // 512 MMAs per repetition
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaMemset(d_acc, 0, acc_n * sizeof(float));
    kern<<<1, 128, 200 * 1024>>>(bmn, n_groups, d_init, d_t, d_acc);
    long long h[8];
    cudaError_t e = cudaMemcpy(h, d_t, 48, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("G=%d %s: CUDA error %s\n", G, names[v], cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(v == 0 ? ref : got, d_acc, acc_n * sizeof(float), cudaMemcpyDeviceToHost);
    size_t bad = 0;
    double sum = 0;
    if (v > 0) for (size_t i = 0; i < (size_t)G * 128 * 64; ++i) bad += (got[i] != ref[i]);
    for (size_t i = 0; i < (size_t)G * 128 * 64; ++i) sum += (v == 0 ? ref : got)[i];
    printf("G=%d  %-24s issue %6lld  total %6lld cyc  -> %5.1f cyc/MMA   mismatches vs plain: %zu  (checksum %.1f)\n", G, names[v], h[4], h[5],
           (double)h[5] / 512.0, bad, sum);
    fflush(stdout);
  };
  for (bmn = 0; bmn < 2; ++bmn) {
  printf("---- B operand %s\n", bmn ? "MN-major (as in the recurrent kernel)" : "K-major");
  run(1, 0, bench<1, 0>); run(1, 1, bench<1, 1>);
  run(2, 0, bench<2, 0>); run(2, 1, bench<2, 1>); run(2, 2, bench<2, 2>);
  run(4, 0, bench<4, 0>); run(4, 1, bench<4, 1>); run(4, 2, bench<4, 2>);
  run(8, 0, bench<8, 0>); run(8, 1, bench<8, 1>); run(8, 2, bench<8, 2>);
  }
  return 0;
}
