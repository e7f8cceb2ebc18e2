// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16) as a function of N, operand layout
// (SWIZZLE_NONE K-major / MN-major, stride choices) and accumulator chaining.  One CTA, one issuing thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_umma ubench_umma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void umma_lh(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n" ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_elect(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t.reg .b32 rx;\n\telect.sync rx|q, 0xffffffff;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
               "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n" ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t dlo(uint32_t a, uint32_t lbo) { return ((a >> 4) & 0x3FFF) | ((lbo >> 4) << 16); }
__device__ __forceinline__ uint32_t dhi(uint32_t sbo, uint32_t swz) { return ((sbo >> 4) & 0x3FFF) | (1u << 14) | (swz << 29); }

struct Cfg { int N, n_mma, n_acc, a_sbo, a_kstep, b_mn, b_lbo, b_sbo, b_kstep, a_swz, b_swz, a_lbo, uniform, M = 128; };

__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
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
  if (threadIdx.x == 0 && !c.uniform) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)c.b_mn << 16) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 128 * 1024;
    const uint32_t ahi = dhi(c.a_sbo, c.a_swz), bhi = dhi(c.b_sbo, c.b_swz);
    uint32_t parity = 0;
    for (int rep = 0; rep < 3; ++rep) {
      uint32_t alo = dlo(a0, c.a_lbo), blo = dlo(b0, c.b_lbo);
      const long long t0 = clock64();
      for (int i = 0; i < c.n_mma; ++i) {
        const uint32_t d = tmem + (uint32_t)(i % c.n_acc) * (uint32_t)c.N;
        umma_lh(d, alo, ahi, blo, bhi, idesc, i >= c.n_acc);
        alo += c.a_kstep >> 4;
        blo += c.b_kstep >> 4;
      }
      const long long t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      { long long tw = clock64(); while (!try_wait(smem_u32(&bar), parity)) { if (clock64() - tw > 200000000LL) __trap(); } }
      parity ^= 1;
      const long long t2 = clock64();
      out[rep * 2] = t1 - t0;
      out[rep * 2 + 1] = t2 - t0;
    }
  }
  if (threadIdx.x < 32 && c.uniform == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)c.b_mn << 16) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 128 * 1024;
    const uint32_t ahi = dhi(c.a_sbo, c.a_swz), bhi = dhi(c.b_sbo, c.b_swz);
    uint32_t parity = 0;
    for (int rep = 0; rep < 3; ++rep) {
      uint32_t alo = dlo(a0, c.a_lbo), blo = dlo(b0, c.b_lbo);
      const long long t0 = clock64();
      for (int i = 0; i < c.n_mma; ++i) {
        const uint32_t d = tmem + (uint32_t)(i % c.n_acc) * (uint32_t)c.N;
        umma_elect(d, alo, ahi, blo, bhi, idesc, i >= c.n_acc);
        alo += c.a_kstep >> 4;
        blo += c.b_kstep >> 4;
      }
      const long long t1 = clock64();
      if (threadIdx.x == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      { long long tw = clock64(); while (!try_wait(smem_u32(&bar), parity)) { if (clock64() - tw > 200000000LL) __trap(); } }
      parity ^= 1;
      const long long t2 = clock64();
      if (threadIdx.x == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
    }
  }
  if (threadIdx.x < 32 && c.uniform >= 2) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)c.b_mn << 16) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 128 * 1024;
    const uint32_t ahi = dhi(c.a_sbo, c.a_swz), bhi = dhi(c.b_sbo, c.b_swz);
    uint32_t parity = 0;
    uint32_t alo[16], blo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { alo[i] = dlo(a0, c.a_lbo) + (uint32_t)i * (c.a_kstep >> 4); blo[i] = dlo(b0, c.b_lbo) + (uint32_t)i * (c.b_kstep >> 4); }
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      if (c.uniform == 2) {
#pragma unroll
        for (int i = 0; i < 16; ++i) umma_elect(tmem, alo[i], ahi, blo[i], bhi, idesc, i > 0);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) umma_elect(tmem + (uint32_t)(i & 3) * 32u, alo[i], ahi, blo[i], bhi, idesc, i > 3);
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
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Named { const char* name; Cfg c; };
  // A K-major no-swizzle: LBO=128 (k chunk), SBO=a_sbo, K step 256 B.  B MN-major no-swizzle: LBO=b_lbo (k-group), SBO=128 (n-group)
  // b K-major no-swizzle: LBO=128, SBO=b_sbo, step 256.
  Named cfgs[] = {
      {"N32  thread0 chain1", {32, 16, 1, 4096, 256, 1, 512, 128, 1024, 0, 0, 128, 0}},
      {"N32  uniform chain1", {32, 16, 1, 4096, 256, 1, 512, 128, 1024, 0, 0, 128, 1}},
      {"N32  precomputed descs, chain1", {32, 16, 1, 4096, 256, 1, 512, 128, 1024, 0, 0, 128, 2}},
      {"N32  precomputed descs, 4 accumulators", {32, 16, 4, 4096, 256, 1, 512, 128, 1024, 0, 0, 128, 3}},
      {"N64  precomputed descs, chain1", {64, 16, 1, 4096, 256, 1, 1024, 128, 2048, 0, 0, 128, 2}},
      {"N128 precomputed descs, chain1", {128, 16, 1, 4096, 256, 1, 2048, 128, 0, 0, 0, 128, 2}},
      {"N256 precomputed descs, chain1", {256, 16, 1, 4096, 256, 1, 4096, 128, 0, 0, 0, 128, 2}},
      {"N32  uniform chain4", {32, 16, 4, 4096, 256, 1, 512, 128, 1024, 0, 0, 128, 1}},
      {"N64  uniform chain1", {64, 16, 1, 4096, 256, 1, 1024, 128, 2048, 0, 0, 128, 1}},
      {"N128 uniform chain1", {128, 16, 1, 4096, 256, 1, 2048, 128, 0, 0, 0, 128, 1}},
      {"N256 uniform chain1", {256, 16, 1, 4096, 256, 1, 4096, 128, 0, 0, 0, 128, 1}},
      {"N256 uniform chain2", {256, 16, 2, 4096, 256, 1, 4096, 128, 0, 0, 0, 128, 1}},
      {"N32  uniform n=64 chain1 (same operands)", {32, 64, 1, 4096, 0, 1, 512, 128, 0, 0, 0, 128, 1}},
      {"N32  uniform n=64 chain8 (same operands)", {32, 64, 8, 4096, 0, 1, 512, 128, 0, 0, 0, 128, 1}},
      {"N128 uniform n=64 chain1 (same operands)", {128, 64, 1, 4096, 0, 1, 2048, 128, 0, 0, 0, 128, 1}},
      {"N256 uniform n=64 chain1 (same operands)", {256, 64, 1, 4096, 0, 1, 4096, 128, 0, 0, 0, 128, 1}},
      {"N32  uniform n=64 B kmajor", {32, 64, 1, 4096, 0, 0, 128, 4096, 0, 0, 0, 128, 1}},
      {"M64 N32  precomputed descs, chain1", {32, 16, 1, 4096, 256, 1, 512, 128, 1024, 0, 0, 128, 2, 64}},
      {"M64 N32  precomputed descs, 4 acc", {32, 16, 4, 4096, 256, 1, 512, 128, 1024, 0, 0, 128, 3, 64}},
      {"M64 N64  precomputed descs, chain1", {64, 16, 1, 4096, 256, 1, 1024, 128, 2048, 0, 0, 128, 2, 64}},
      {"M128 N16 precomputed descs, chain1", {16, 16, 1, 4096, 256, 1, 256, 128, 512, 0, 0, 128, 2, 128}},
  };
  for (auto& nc : cfgs) {
    bench<<<1, 128, 200 * 1024>>>(nc.c, d);
    long long h[8];
    cudaError_t e = cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", nc.name, cudaGetErrorString(e)); return 1; }
    printf("%-46s n=%2d  issue %6lld  total %6lld cyc  -> %.1f cyc/MMA (rep2: %lld)\n", nc.name, nc.c.n_mma, h[2], h[3], (double)h[3] / nc.c.n_mma, h[5]);
    fflush(stdout);
  }
  return 0;
}
