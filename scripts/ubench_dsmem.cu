// Microbenchmark: distributed-shared-memory exchange cost inside a thread-block cluster on sm_100a.
// Every CTA of a cluster of C pushes `bytes` to each of its C-1 peers per round, either with
//   mode 0: cp.async.bulk.shared::cluster.shared::cta (bulk copy, completes on the PEER's mbarrier), or
//   mode 1: st.shared::cluster.v4 from 128 threads followed by a cluster barrier,
// and rounds are separated by a cluster barrier.  Reports cycles per round (bytes=0 gives the sync floor).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_dsmem ubench_dsmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}

struct Cfg { int C, bytes, mode, rounds; };

__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];   // [0,64K) send buffer, [64K, 64K + 8*16K...) receive slots
  __shared__ uint64_t bar;
  const uint32_t me = cluster_rank();
  const uint32_t send = smem_u32(smem), recv = send + 65536u;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  cluster_sync();
  uint32_t parity = 0;
  long long t_first = 0, t_total = 0;
  for (int rnd = 0; rnd < c.rounds + 1; ++rnd) {
    const long long t0 = clock64();
    if (c.mode == 0) {
      if (threadIdx.x == 0) {
        if (c.bytes > 0) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)(c.bytes * (c.C - 1))) : "memory");
          for (int p = 1; p < c.C; ++p) {
            const uint32_t peer = (me + p) % c.C;
            const uint32_t dst = mapa(recv + (uint32_t)(me * c.bytes), peer);   // slot indexed by the sender
            const uint32_t pbar = mapa(smem_u32(&bar), peer);
            asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "r"(send), "r"((uint32_t)c.bytes), "r"(pbar)
                         : "memory");
          }
          long long tw = clock64();
          while (!try_wait(smem_u32(&bar), parity)) {
            if (clock64() - tw > 400000000LL) __trap();
          }
          parity ^= 1;
        }
      }
      __syncthreads();
    } else {
      const int n16 = c.bytes / 16;
      for (int p = 1; p < c.C; ++p) {
        const uint32_t peer = (me + p) % c.C;
        const uint32_t dst = mapa(recv + (uint32_t)(me * c.bytes), peer);
        for (int i = threadIdx.x; i < n16; i += blockDim.x) {
          const uint4 v = reinterpret_cast<const uint4*>(smem)[i];
          asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16u * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
      }
    }
    cluster_sync();
    const long long t1 = clock64();
    if (rnd == 0) t_first = t1 - t0; else t_total += t1 - t0;
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    out[0] = t_first;
    out[1] = t_total / c.rounds;
  }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  const int smem_bytes = 65536 + 8 * 16384 + 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(bench, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int Cs[] = {2, 4, 8};
  const int sizes[] = {0, 2048, 8192, 16384};
  for (int mode = 0; mode < 2; ++mode)
    for (int C : Cs)
      for (int bytes : sizes) {
        for (int grid_clusters : {1, 16}) {
          Cfg c{C, bytes, mode, 32};
          cudaLaunchConfig_t cfg{};
          cfg.gridDim = dim3(C * grid_clusters);
          cfg.blockDim = dim3(128);
          cfg.dynamicSmemBytes = smem_bytes;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension;
          at[0].val.clusterDim.x = C;
          at[0].val.clusterDim.y = 1;
          at[0].val.clusterDim.z = 1;
          cfg.attrs = at;
          cfg.numAttrs = 1;
          cudaError_t e = cudaLaunchKernelEx(&cfg, bench, c, d_out);
          if (e == cudaSuccess) e = cudaDeviceSynchronize();
          long long h[2] = {0, 0};
          cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) {
            printf("mode %d C=%d bytes=%d clusters=%d: ERROR %s\n", mode, C, bytes, grid_clusters, cudaGetErrorString(e));
            return 1;
          }
          const double out_bytes = (double)bytes * (C - 1);
          printf("mode %s C=%d clusters=%2d bytes/peer=%5d: %6lld cyc/round (first %lld)  -> %.1f B/clk out per CTA\n", mode ? "st.v4 " : "bulk  ", C,
                 grid_clusters, bytes, h[1], h[0], out_bytes / (double)h[1]);
        }
      }
  return 0;
}
