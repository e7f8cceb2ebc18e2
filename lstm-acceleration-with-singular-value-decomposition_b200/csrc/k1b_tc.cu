// K1b: tcgen05 BF16 tensor-core persistent recurrent kernel (reduced-precision engine).
//
// One CTA owns a tile of N=32 sequences for all T steps of ONE layer; layers run back to back,
// handing the bf16 hidden sequence over through HBM in the exact shared-memory operand image the
// next layer's MMAs consume (one 1-D bulk copy per step).  Per step the low-rank cell is two
// dependent thin contractions on the 5th-gen tensor cores, SWAP-AB so that the weight dimension fills
// the 128-row MMA and the batch tile is the (small) N:
//   S1u: t_u[r_u x N]  = (L_u sigma_u)^T [r_u x H]   . h(t-1) [H x N]
//   S1w: t_w[r_w x N]  = (L_w sigma_w)^T [r_w x Hin] . in(t)  [Hin x N]      (layers >= 1)
//   S2 : z [4H x N]    = [R_u ; R_w]^T   [4H x K2]   . [t_u ; t_w | x(t)]    (layer 0: x enters S2 directly
//                                                                             through the dense 16x4H W)
// Accumulators live in TMEM (S1: 2 x 32 columns; S2: two buffers of 4 gates x 32 columns, so the
// gate/cell epilogue of unit block ub overlaps the MMAs of block ub+1).  Weight factors are packed
// once (bf16, K-major core-matrix images) and stay resident in shared memory for the whole launch.
// Warp roles: warp 0 = bulk-copy producer (input prefetch ring + output stores), warp 1 = MMA issuer
// (one elected thread), warps 2-5 = epilogue (TMEM -> registers -> activations -> bf16 operand in smem).
// Cell state c stays in FP32 registers for all T steps.
//
// Replaces SingularLSTMCell.call / ReducedLSTMCell.call + backend.rnn for large batches
// (reference code/svd_classes_v3.py:116-145, 317-328, 405-434) at BF16 precision; the FP32 engines
// remain the parity path.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace svdlstm {

namespace {

constexpr int kN = 32;            // sequences per CTA (MMA N)
constexpr int kInStages = 3;      // max input prefetch ring depth
constexpr int kTcThreads = 192;   // 6 warps

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped launch (CUDA error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FFu) == 0 && clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]   (kind::f16: bf16 x bf16 -> fp32)
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, descriptors given as (lo, hi) words: per-MMA descriptor update is ONE 32-bit add on the lo word
__device__ __forceinline__ void umma_lh(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_approx(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ------------------------------------------------------------------------------------------------
// descriptors (cute/arch/mma_sm100_desc.hpp bit layout; SWIZZLE_NONE canonical layouts)
//   K-major : elem(mn,k) at (mn%8)*16 + (mn/8)*SBO + (k%8)*2 + (k/8)*LBO
//   MN-major: elem(mn,k) at (mn%8)*2  + (mn/8)*SBO + (k%8)*16 + (k/8)*LBO
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version 1 (Blackwell)
  return d;                 // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (0)
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                      // c_format  F32
         | (1u << 7)                    // a_format  BF16
         | (1u << 10)                   // b_format  BF16
         | ((uint32_t)a_mn_major << 15) // a_major
         | ((uint32_t)b_mn_major << 16) // b_major
         | ((uint32_t)(N >> 3) << 17)   // n_dim
         | ((uint32_t)(M >> 4) << 24);  // m_dim
}

// MN-major activation tile image [K x N=32]: k-groups of 8 are 512 B apart, n-groups of 8 128 B apart.
constexpr uint32_t kActLBO = 512, kActSBO = 128;
__host__ __device__ inline uint32_t act_tile_bytes(int K) { return (uint32_t)(K / 8) * 512u; }
__host__ __device__ inline uint32_t act_offset(int k, int n) {
  return (uint32_t)(k / 8) * 512u + (uint32_t)(n / 8) * 128u + (uint32_t)(k % 8) * 16u + (uint32_t)(n % 8) * 2u;
}

// ------------------------------------------------------------------------------------------------
// per-layer launch parameters
// ------------------------------------------------------------------------------------------------
struct TcLayerParams {
  const uint8_t* a1u;     // (L_u sigma_u)^T  K-major image, rows r_u_pad8, K = H
  const uint8_t* a1w;     // (L_w sigma_w)^T  K-major image, rows r_w_pad8, K = Hin   (nullptr for layer 0)
  const uint8_t* a2;      // [R_u ; R_w|W0]^T K-major image, 4H rows in (unit-block, gate) tiles of 128, K = K2
  const float* bias;      // [4H] in tile order
  const uint8_t* in_seq;  // activation tile images [cta][t], K = Kin
  uint8_t* out_seq;       // activation tile images [cta][t], K = H
  int H, Kin, T;
  int ru, rw;             // true ranks
  int ru_pad, rin_pad;    // multiples of 16: K extents of the two parts of the S2 contraction
  int has_s1w;            // layers >= 1
  int in_stages;          // depth of the input prefetch ring (2 or 3)
  uint32_t a1u_bytes, a1w_bytes, a2_bytes;
  long long* dbg;         // optional timeline buffer (CTA 0): 16 clock64 stamps per step; nullptr = off
};

struct TcSmemPlan {
  uint32_t a1u, a1w, a2, hbuf, tbuf, inbuf, bars, tmem_slot, total;
};

__host__ __device__ inline TcSmemPlan tc_plan(const TcLayerParams& p) {
  TcSmemPlan s;
  uint32_t off = 0;
  s.a1u = off; off += (p.a1u_bytes + 127u) & ~127u;
  s.a1w = off; off += (p.a1w_bytes + 127u) & ~127u;
  s.a2 = off;  off += (p.a2_bytes + 127u) & ~127u;
  s.hbuf = off; off += act_tile_bytes(p.H);
  s.tbuf = off; off += act_tile_bytes(p.ru_pad + (p.has_s1w ? p.rin_pad : 0));
  s.inbuf = off; off += (uint32_t)p.in_stages * act_tile_bytes(p.Kin);
  s.bars = off; off += 128;
  s.tmem_slot = off; off += 16;
  s.total = off + 2048;   // tail guard: 128-row MMA tiles over-read past short (r < 128 row) images
  return s;
}

// barrier slots (8 B each) inside the `bars` block
enum { BAR_IN_FULL = 0, BAR_IN_EMPTY = kInStages, BAR_S1_FULL = 2 * kInStages, BAR_T_READY, BAR_S2_FULL0, BAR_S2_FULL1,
       BAR_S2_EMPTY0, BAR_S2_EMPTY1, BAR_H_READY, BAR_H_FREE, BAR_COUNT };
static_assert(BAR_COUNT * 8 <= 128, "barrier block too small");

#define TC_STAMP(slot) do { if (p.dbg != nullptr && blockIdx.x == 0 && t < 64) p.dbg[t * 16 + (slot)] = clock64(); } while (0)

__global__ void __launch_bounds__(kTcThreads, 1) lstm_tc_layer_kernel(const TcLayerParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const TcSmemPlan sp = tc_plan(p);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  const int H = p.H, T = p.T;
  const int nub = H / 128;
  const int nst = p.in_stages;
  const uint32_t in_tile = act_tile_bytes(p.Kin), h_tile = act_tile_bytes(H);
  const uint32_t bar0 = sbase + sp.bars;
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  // ---- one-time setup ---------------------------------------------------------------------------
  // zero the activation buffers (h(-1) = 0, padded t rows must be finite)
  for (uint32_t i = threadIdx.x * 16u; i < sp.inbuf - sp.hbuf; i += kTcThreads * 16u)
    *reinterpret_cast<uint4*>(smem + sp.hbuf + i) = make_uint4(0, 0, 0, 0);
  for (uint32_t i = threadIdx.x * 16u; i < 2048u; i += kTcThreads * 16u)
    *reinterpret_cast<uint4*>(smem + sp.total - 2048u + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kInStages; ++s) {
      mbar_init(bar(BAR_IN_FULL + s), 1);
      mbar_init(bar(BAR_IN_EMPTY + s), 1);
    }
    mbar_init(bar(BAR_S1_FULL), 1);
    mbar_init(bar(BAR_T_READY), 128);
    mbar_init(bar(BAR_S2_FULL0), 1);
    mbar_init(bar(BAR_S2_FULL1), 1);
    mbar_init(bar(BAR_S2_EMPTY0), 128);
    mbar_init(bar(BAR_S2_EMPTY1), 128);
    mbar_init(bar(BAR_H_READY), 128);
    mbar_init(bar(BAR_H_FREE), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(sbase + sp.tmem_slot, 512);
  fence_proxy_async();   // generic-proxy zero fill -> visible to the async proxy (MMA / bulk copies)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + sp.tmem_slot);
  // TMEM columns: [0,32) S1u, [32,64) S1w, [64,192) S2 buffer 0 (4 gates x 32), [192,320) S2 buffer 1
  const uint32_t tm_s1u = tmem, tm_s1w = tmem + 32, tm_s2 = tmem + 64;

  if (warp == 0) {
    // ======================= producer: weights once, then the input ring + output stores ==========
    if (lane == 0) {
      // weights: one transaction barrier (reuse IN_FULL[0] phase 0 would complicate parities; use S2_FULL1's
      // sibling-free slot instead: H_FREE is idle until the first h store) -> dedicated wait below
      const uint32_t wbytes = p.a1u_bytes + p.a1w_bytes + p.a2_bytes;
      mbar_expect_tx(bar(BAR_H_FREE), wbytes);
      // bulk copies are limited in size per instruction; chunk at 64 KB
      auto copy_big = [&](uint32_t dst, const uint8_t* src, uint32_t bytes) {
        for (uint32_t o = 0; o < bytes; o += 65536u) {
          const uint32_t n = bytes - o < 65536u ? bytes - o : 65536u;
          bulk_g2s(sbase + dst + o, src + o, n, bar(BAR_H_FREE));
        }
      };
      copy_big(sp.a1u, p.a1u, p.a1u_bytes);
      if (p.has_s1w) copy_big(sp.a1w, p.a1w, p.a1w_bytes);
      copy_big(sp.a2, p.a2, p.a2_bytes);
      // input ring + output stores
      const uint8_t* src = p.in_seq + (size_t)cta * T * in_tile;
      uint8_t* out = p.out_seq + (size_t)cta * T * h_tile;
      uint32_t ph_empty = 0;   // per-stage parity bits
      auto load_step = [&](int tl) {
        const int s = tl % nst;
        if (tl >= nst) {
          mbar_wait(bar(BAR_IN_EMPTY + s), (ph_empty >> s) & 1u);
          ph_empty ^= 1u << s;
        }
        mbar_expect_tx(bar(BAR_IN_FULL + s), in_tile);
        bulk_g2s(sbase + sp.inbuf + s * in_tile, src + (size_t)tl * in_tile, in_tile, bar(BAR_IN_FULL + s));
      };
      for (int tl = 0; tl < nst - 1 && tl < T; ++tl) load_step(tl);
      for (int t = 0; t < T; ++t) {
        if (t + nst - 1 < T) load_step(t + nst - 1);
        // h(t) complete in smem -> ship it to HBM, then let the epilogue overwrite the buffer
        mbar_wait(bar(BAR_H_READY), (uint32_t)(t & 1));
        bulk_s2g(out + (size_t)t * h_tile, sbase + sp.hbuf, h_tile);
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(bar(BAR_H_FREE));
      }
      bulk_wait_all0();
    }
  } else if (warp == 1) {
    // ======================= MMA issuer ==========================================================
    if (lane == 0) {
      const uint32_t idesc_mn = make_idesc(128, kN, 0, 1);   // A K-major (weights), B MN-major (activations)
      // weights landed?
      mbar_wait(bar(BAR_H_FREE), 0);
      uint32_t ph_in = 0, ph_t = 0, ph_hready = 0, ph_s2e = 0;
      const int K2 = p.ru_pad + p.rin_pad;
      // descriptor words: hi = SBO | version, lo = start address | LBO.  K advances by 16 elements per MMA:
      // +256 B on K-major weight images (lo += 16), +1024 B on MN-major activation tiles (lo += 64).
      const uint32_t act_hi = desc_hi(kActSBO);
      const uint32_t a1u_hi = desc_hi((uint32_t)H * 16u), a1u_lo0 = desc_lo(sbase + sp.a1u, 128u);
      const uint32_t a1w_hi = desc_hi((uint32_t)p.Kin * 16u), a1w_lo0 = desc_lo(sbase + sp.a1w, 128u);
      const uint32_t a2_hi = desc_hi((uint32_t)K2 * 16u), a2_lo0 = desc_lo(sbase + sp.a2, 128u);
      const uint32_t a2_tile_lo = (128u * (uint32_t)K2 * 2u) >> 4;
      const uint32_t h_lo0 = desc_lo(sbase + sp.hbuf, kActLBO), t_lo0 = desc_lo(sbase + sp.tbuf, kActLBO);
      const uint32_t in_lo0 = desc_lo(sbase + sp.inbuf, kActLBO), in_stage_lo = in_tile >> 4;
      const int n_s1u = H / 16, n_s1w = p.has_s1w ? p.Kin / 16 : 0;
      const int n_s2t = p.has_s1w ? K2 / 16 : p.ru_pad / 16;     // K slices of S2 read from the t buffer
      const int n_s2x = p.has_s1w ? 0 : p.rin_pad / 16;          // K slices of S2 read from the x stage (layer 0)
      for (int t = 0; t < T; ++t) {
        const int s = t % nst;
        if (t > 0) {
          mbar_wait(bar(BAR_H_READY), ph_hready);
          ph_hready ^= 1u;
        }
        TC_STAMP(0);   // MMA: h(t-1) ready seen
        mbar_wait(bar(BAR_IN_FULL + s), (ph_in >> s) & 1u);
        ph_in ^= 1u << s;
        tc_fence_after();
        TC_STAMP(1);   // MMA: in(t) ready seen
        const uint32_t in_lo = in_lo0 + (uint32_t)s * in_stage_lo;
        // ---- S1u: t_u = A1u . h(t-1)    K = H
        {
          uint32_t alo = a1u_lo0, blo = h_lo0;
#pragma unroll 8
          for (int kk = 0; kk < n_s1u; ++kk) {
            umma_lh(tm_s1u, alo, a1u_hi, blo, act_hi, idesc_mn, kk > 0);
            alo += 16u;
            blo += 64u;
          }
        }
        // ---- S1w: t_w = A1w . in(t)     K = Kin
        {
          uint32_t alo = a1w_lo0, blo = in_lo;
#pragma unroll 8
          for (int kk = 0; kk < n_s1w; ++kk) {
            umma_lh(tm_s1w, alo, a1w_hi, blo, act_hi, idesc_mn, kk > 0);
            alo += 16u;
            blo += 64u;
          }
        }
        umma_commit(bar(BAR_S1_FULL));
        if (p.has_s1w) umma_commit(bar(BAR_IN_EMPTY + s));   // in(t) consumed once S1w completes
        TC_STAMP(2);   // MMA: S1 issued + committed
        // ---- S2: z = A2 . [t_u ; t_w | x(t)]
        mbar_wait(bar(BAR_T_READY), ph_t);
        ph_t ^= 1u;
        tc_fence_after();
        TC_STAMP(3);   // MMA: t operand ready seen
        uint32_t tile_lo = a2_lo0;
        for (int ub = 0; ub < nub; ++ub) {
          const int buf = ub & 1;
          // wait until the epilogue drained this TMEM buffer (first use of each buffer passes: parity trick)
          mbar_wait(bar(BAR_S2_EMPTY0 + buf), ((ph_s2e >> buf) & 1u) ^ 1u);
          ph_s2e ^= 1u << buf;
          tc_fence_after();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t d = tm_s2 + (uint32_t)buf * 128u + (uint32_t)g * 32u;
            uint32_t alo = tile_lo, blo = t_lo0;
#pragma unroll 4
            for (int kk = 0; kk < n_s2t; ++kk) {
              umma_lh(d, alo, a2_hi, blo, act_hi, idesc_mn, kk > 0);
              alo += 16u;
              blo += 64u;
            }
            blo = in_lo;
            for (int kk = 0; kk < n_s2x; ++kk) {
              umma_lh(d, alo, a2_hi, blo, act_hi, idesc_mn, 1u);
              alo += 16u;
              blo += 64u;
            }
            tile_lo += a2_tile_lo;
          }
          umma_commit(bar(BAR_S2_FULL0 + buf));
          TC_STAMP(4 + ub);   // MMA: S2 block ub issued + committed
        }
        if (!p.has_s1w) umma_commit(bar(BAR_IN_EMPTY + s));   // layer 0: x(t) is consumed by S2
      }
    }
  } else {
    // ======================= epilogue warps (128 threads; thread = TMEM lane = one row) ==============
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;             // row of every 128-row tile handled by this thread
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float cst[2][kN];                          // cell state of unit (ub*128+row), FP32, all T steps  (H <= 256)
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int n = 0; n < kN; ++n) cst[u][n] = 0.f;
    float bi[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int g = 0; g < 4; ++g) bi[u][g] = (u < nub) ? p.bias[(u * 4 + g) * 128 + row] : 0.f;
    uint32_t ph_s1 = 0, ph_s2f = 0, ph_hfree = 1;   // H_FREE phase 0 was consumed by the weight load
    for (int t = 0; t < T; ++t) {
      // ---- epilogue 1: t_u / t_w accumulators -> bf16 rows of the S2 B operand ---------------------
      mbar_wait(bar(BAR_S1_FULL), ph_s1);
      ph_s1 ^= 1u;
      tc_fence_after();
      if (threadIdx.x == 64) TC_STAMP(8);   // EPI: S1 accumulators seen
      {
        const bool do_u = (q * 32) < p.ru_pad;       // warp-uniform: any valid row in this quarter?
        const bool do_w = p.has_s1w && (q * 32) < p.rin_pad;
        uint32_t r[16];
        if (do_u) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            tmem_ld16(tm_s1u + lane_addr + half * 16, r);
            tmem_ld_wait();
            if (row < p.ru_pad) {
              const bool live = row < p.ru;
              uint4 v0, v1;
              v0.x = live ? pack_bf16(__uint_as_float(r[0]), __uint_as_float(r[1])) : 0u;
              v0.y = live ? pack_bf16(__uint_as_float(r[2]), __uint_as_float(r[3])) : 0u;
              v0.z = live ? pack_bf16(__uint_as_float(r[4]), __uint_as_float(r[5])) : 0u;
              v0.w = live ? pack_bf16(__uint_as_float(r[6]), __uint_as_float(r[7])) : 0u;
              v1.x = live ? pack_bf16(__uint_as_float(r[8]), __uint_as_float(r[9])) : 0u;
              v1.y = live ? pack_bf16(__uint_as_float(r[10]), __uint_as_float(r[11])) : 0u;
              v1.z = live ? pack_bf16(__uint_as_float(r[12]), __uint_as_float(r[13])) : 0u;
              v1.w = live ? pack_bf16(__uint_as_float(r[14]), __uint_as_float(r[15])) : 0u;
              uint8_t* dst = smem + sp.tbuf + act_offset(row, half * 16);
              *reinterpret_cast<uint4*>(dst) = v0;
              *reinterpret_cast<uint4*>(dst + kActSBO) = v1;
            }
          }
        }
        if (do_w) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            tmem_ld16(tm_s1w + lane_addr + half * 16, r);
            tmem_ld_wait();
            if (row < p.rin_pad) {
              const bool live = row < p.rw;
              uint4 v0, v1;
              v0.x = live ? pack_bf16(__uint_as_float(r[0]), __uint_as_float(r[1])) : 0u;
              v0.y = live ? pack_bf16(__uint_as_float(r[2]), __uint_as_float(r[3])) : 0u;
              v0.z = live ? pack_bf16(__uint_as_float(r[4]), __uint_as_float(r[5])) : 0u;
              v0.w = live ? pack_bf16(__uint_as_float(r[6]), __uint_as_float(r[7])) : 0u;
              v1.x = live ? pack_bf16(__uint_as_float(r[8]), __uint_as_float(r[9])) : 0u;
              v1.y = live ? pack_bf16(__uint_as_float(r[10]), __uint_as_float(r[11])) : 0u;
              v1.z = live ? pack_bf16(__uint_as_float(r[12]), __uint_as_float(r[13])) : 0u;
              v1.w = live ? pack_bf16(__uint_as_float(r[14]), __uint_as_float(r[15])) : 0u;
              uint8_t* dst = smem + sp.tbuf + act_offset(p.ru_pad + row, half * 16);
              *reinterpret_cast<uint4*>(dst) = v0;
              *reinterpret_cast<uint4*>(dst + kActSBO) = v1;
            }
          }
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(bar(BAR_T_READY));
      if (threadIdx.x == 64) TC_STAMP(9);   // EPI: t operand written + arrived

      // ---- epilogue 2: gates + cell update per unit block ------------------------------------------
      // h(t-1) tile must have been copied out before it is overwritten
      if (t > 0) {
        mbar_wait(bar(BAR_H_FREE), ph_hfree);
        ph_hfree ^= 1u;
      }
#pragma unroll
      for (int ub = 0; ub < 2; ++ub) {
        if (ub < nub) {
          const int buf = ub & 1;
          mbar_wait(bar(BAR_S2_FULL0 + buf), (ph_s2f >> buf) & 1u);
          ph_s2f ^= 1u << buf;
          tc_fence_after();
          if (threadIdx.x == 64) TC_STAMP(10 + 2 * ub);   // EPI: z block ub seen
          const uint32_t tb = tm_s2 + (uint32_t)buf * 128u + lane_addr;
          const int unit = ub * 128 + row;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t zi[16], zf[16], zg[16], zo[16];
            tmem_ld16(tb + 0 * 32 + half * 16, zi);
            tmem_ld16(tb + 1 * 32 + half * 16, zf);
            tmem_ld16(tb + 2 * 32 + half * 16, zg);
            tmem_ld16(tb + 3 * 32 + half * 16, zo);
            tmem_ld_wait();
            float hv[16];
#pragma unroll
            for (int n = 0; n < 16; ++n) {
              const float ig = sigmoid_approx(__uint_as_float(zi[n]) + bi[ub][0]);
              const float fg = sigmoid_approx(__uint_as_float(zf[n]) + bi[ub][1]);
              const float gg = tanh_approx(__uint_as_float(zg[n]) + bi[ub][2]);
              const float og = sigmoid_approx(__uint_as_float(zo[n]) + bi[ub][3]);
              const float c = fmaf(fg, cst[ub][half * 16 + n], ig * gg);
              cst[ub][half * 16 + n] = c;
              hv[n] = og * tanh_approx(c);
            }
            uint4 v0, v1;
            v0.x = pack_bf16(hv[0], hv[1]);   v0.y = pack_bf16(hv[2], hv[3]);
            v0.z = pack_bf16(hv[4], hv[5]);   v0.w = pack_bf16(hv[6], hv[7]);
            v1.x = pack_bf16(hv[8], hv[9]);   v1.y = pack_bf16(hv[10], hv[11]);
            v1.z = pack_bf16(hv[12], hv[13]); v1.w = pack_bf16(hv[14], hv[15]);
            uint8_t* dst = smem + sp.hbuf + act_offset(unit, half * 16);
            *reinterpret_cast<uint4*>(dst) = v0;
            *reinterpret_cast<uint4*>(dst + kActSBO) = v1;
          }
          tc_fence_before();
          mbar_arrive(bar(BAR_S2_EMPTY0 + buf));
          if (threadIdx.x == 64) TC_STAMP(11 + 2 * ub);   // EPI: block ub done
        }
      }
      fence_proxy_async();
      mbar_arrive(bar(BAR_H_READY));
      if (threadIdx.x == 64) TC_STAMP(14);      // EPI: h(t) published
    }
  }

  // ---- teardown -----------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// packing / layout kernels (run once per weight update, or once per forward for the sequences)
// ------------------------------------------------------------------------------------------------
// K-major weight image: elem(row,k) at (row/8)*(K*16) + (k/8)*128 + (row%8)*16 + (k%8)*2
__device__ __forceinline__ size_t kmaj_off(int row, int k, int K) {
  return (size_t)(row / 8) * ((size_t)K * 16) + (size_t)(k / 8) * 128 + (size_t)(row % 8) * 16 + (size_t)(k % 8) * 2;
}

// A1 image: rows = rank index j (padded to 8), K = kin: value = left[k*ld + j] * scale[j]
__global__ void pack_a1_kernel(const float* __restrict__ left, int ld, const float* __restrict__ scale, int r, int r_pad8, int K,
                               __nv_bfloat16* __restrict__ img) {
  const int total = r_pad8 * K;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int row = idx / K, k = idx - row * K;
    float v = 0.f;
    if (row < r) v = left[(size_t)k * ld + row] * (scale ? scale[row] : 1.f);
    img[kmaj_off(row, k, K) / 2] = __float2bfloat16_rn(v);
  }
}

// effective right-factor element of a block: row kk of the (rank x 4H) matrix, gate column n
__device__ __forceinline__ float block_right(const Block& b, int kk, int n) {
  int rel = n - b.out0;
  if (rel < 0) return 0.f;
  if (b.ident) {
    if (rel < b.rank) return rel == kk ? 1.f : 0.f;
    rel -= b.rank;
  }
  if (rel >= b.ncols) return 0.f;
  return b.right[(size_t)kk * b.right_ld + rel];
}

// A2 image: tile tl = ub*4+g holds gate columns n = g*H + ub*128 + i, i<128; K = ru_pad + rin_pad:
//   k <  ru_pad            : R_u[k][n]                        (0 beyond r_u)
//   k >= ru_pad (layer>=1) : R_w[k-ru_pad][n]                 (0 beyond r_w)
//   k >= ru_pad (layer 0)  : W0[d][n] = sum_j L_w[d][j] s_w[j] R_w[j][n],  d = k-ru_pad < D
__global__ void pack_a2_kernel(Block bw, Block bu, int H, int ru_pad, int rin_pad, int dense_input, int D,
                               __nv_bfloat16* __restrict__ img, float* __restrict__ bias_img, const float* __restrict__ bias) {
  const int K2 = ru_pad + rin_pad;
  const int total = 4 * H * K2;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int rowg = idx / K2, k = idx - rowg * K2;
    const int tl = rowg / 128, i = rowg - tl * 128;
    const int ub = tl / 4, g = tl - ub * 4;
    const int n = g * H + ub * 128 + i;
    float v = 0.f;
    if (k < ru_pad) {
      if (k < bu.rank) v = block_right(bu, k, n);
    } else {
      const int kk = k - ru_pad;
      if (dense_input) {
        if (kk < D) {
          float acc = 0.f;
          for (int j = 0; j < bw.rank; ++j) {
            const float l = bw.left ? bw.left[(size_t)kk * bw.left_ld + j] * (bw.scale ? bw.scale[j] : 1.f) : (kk == j ? 1.f : 0.f);
            acc = fmaf(l, block_right(bw, j, n), acc);
          }
          v = acc;
        }
      } else if (kk < bw.rank) {
        v = block_right(bw, kk, n);
      }
    }
    img[((size_t)tl * 128 * K2 * 2 + kmaj_off(i, k, K2)) / 2] = __float2bfloat16_rn(v);
    if (k == 0) bias_img[rowg] = bias[n];
  }
}

// x (B,T,D) fp32 -> bf16 activation tile images [cta][t] (K = Dpad16)
__global__ void pack_x_kernel(const float* __restrict__ x, int B, int T, int D, int Dpad, uint8_t* __restrict__ img) {
  const size_t total = (size_t)gridDim.y * T * Dpad * kN;   // gridDim.y = number of CTAs (batch tiles)
  const uint32_t tile = act_tile_bytes(Dpad);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (size_t)T * Dpad * kN; idx += (size_t)gridDim.x * blockDim.x) {
    const int cta = blockIdx.y;
    const int t = idx / (Dpad * kN);
    const int rem = idx - (size_t)t * Dpad * kN;
    const int n = rem / Dpad, k = rem - n * Dpad;
    const int b = cta * kN + n;
    float v = 0.f;
    if (b < B && k < D) v = x[((size_t)b * T + t) * D + k];
    *reinterpret_cast<__nv_bfloat16*>(img + ((size_t)cta * T + t) * tile + act_offset(k, n)) = __float2bfloat16_rn(v);
  }
  (void)total;
}

// last layer h tiles -> y (B,T,n_out) = h . dense + bias   (or h itself as fp32 when n_out == 0)
__global__ void unpack_out_kernel(const uint8_t* __restrict__ img, int B, int T, int H, const float* __restrict__ dk,
                                  const float* __restrict__ db, int n_out, float* __restrict__ y) {
  const uint32_t tile = act_tile_bytes(H);
  const int cta = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int item = blockIdx.x * nw + warp; item < T * kN; item += gridDim.x * nw) {
    const int t = item / kN, n = item - t * kN;
    const int b = cta * kN + n;
    if (b >= B) continue;
    const uint8_t* tp = img + ((size_t)cta * T + t) * tile;
    if (n_out > 0) {
      for (int o = 0; o < n_out; ++o) {
        float acc = 0.f;
        for (int k = lane; k < H; k += 32)
          acc = fmaf(__bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(tp + act_offset(k, n))), dk[(size_t)k * n_out + o], acc);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
        if (lane == 0) y[((size_t)b * T + t) * n_out + o] = acc + db[o];
      }
    } else {
      for (int k = lane; k < H; k += 32)
        y[((size_t)b * T + t) * H + k] = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(tp + act_offset(k, n)));
    }
  }
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct TcLayerImg {
  uint8_t *a1u = nullptr, *a1w = nullptr, *a2 = nullptr;
  float* bias = nullptr;
  TcLayerParams prm;
};

struct TcState {
  TcLayerImg layers[kMaxLayers];
  int n_layers = 0;
  uint8_t* seq[2] = {nullptr, nullptr};   // ping-pong activation sequences
  size_t seq_bytes[2] = {0, 0};
  uint8_t* xseq = nullptr;
  size_t xseq_bytes = 0;
};

void tc_free(TcState* s) {
  if (!s) return;
  for (int l = 0; l < kMaxLayers; ++l) {
    if (s->layers[l].a1u) cudaFree(s->layers[l].a1u);
    if (s->layers[l].a1w) cudaFree(s->layers[l].a1w);
    if (s->layers[l].a2) cudaFree(s->layers[l].a2);
    if (s->layers[l].bias) cudaFree(s->layers[l].bias);
  }
  for (int i = 0; i < 2; ++i)
    if (s->seq[i]) cudaFree(s->seq[i]);
  if (s->xseq) cudaFree(s->xseq);
  delete s;
}

static bool tc_layer_params(const ModelDesc& md, int l, TcLayerParams& p, const char** why) {
  const LayerDesc& L = md.layers[l];
  if (L.n_blocks != 2) { *why = "only merged (non-split) cell forms run on the tensor-core engine"; return false; }
  const Block& bw = L.blocks[0];
  const Block& bu = L.blocks[1];
  if (bu.left == nullptr || bw.left == nullptr) { *why = "full (unfactored) cells are not low-rank: use the FP32 engines"; return false; }
  const int H = L.units;
  if (H % 128 != 0 || H > 256) { *why = "tensor-core engine needs units in {128, 256}"; return false; }
  if (bu.rank > 128 || bw.rank > 128) { *why = "ranks above 128 are not supported by the tensor-core engine yet"; return false; }
  p.H = H;
  p.ru = bu.rank;
  p.rw = bw.rank;
  p.ru_pad = round_up(bu.rank, 16);
  p.has_s1w = l > 0;
  if (l == 0) {
    if (L.d_in > 64) { *why = "layer-0 input_dim above 64 is not supported by the tensor-core engine yet"; return false; }
    p.Kin = round_up(L.d_in, 16);
    p.rin_pad = p.Kin;
  } else {
    p.Kin = md.layers[l - 1].units;
    p.rin_pad = round_up(bw.rank, 16);
  }
  p.a1u_bytes = (uint32_t)round_up(p.ru, 8) * H * 2;
  p.a1w_bytes = p.has_s1w ? (uint32_t)round_up(p.rw, 8) * p.Kin * 2 : 0;
  p.a2_bytes = (uint32_t)4 * H * (p.ru_pad + p.rin_pad) * 2;
  p.a1u = p.a1w = p.a2 = nullptr;
  p.bias = nullptr;
  p.dbg = nullptr;
  p.in_seq = nullptr;
  p.out_seq = nullptr;
  p.T = 0;
  p.in_stages = 3;
  if (tc_plan(p).total > 227 * 1024) p.in_stages = 2;
  const TcSmemPlan sp = tc_plan(p);
  if (sp.total > 227 * 1024) { *why = "factor matrices of this rank do not fit the shared memory of one SM (single-CTA engine)"; return false; }
  return true;
}

bool tc_supported(const ModelDesc& md, const ForwardArgs& a, const char** why) {
  if (a.mask || a.h0 || a.h_n || a.c_n) { *why = "mask / initial_state / return_state are FP32-engine features"; return false; }
  if (a.flags & (SVDLSTM_GO_BACKWARDS | SVDLSTM_TIME_MAJOR)) { *why = "go_backwards / time_major are FP32-engine features"; return false; }
  if (!(a.flags & SVDLSTM_RETURN_SEQUENCES)) { *why = "return_sequences=False is an FP32-engine feature"; return false; }
  for (int l = 0; l < md.n_layers; ++l) {
    TcLayerParams p;
    if (!tc_layer_params(md, l, p, why)) return false;
  }
  return true;
}

int run_tc_bf16(const ModelDesc& md, TcState** state, bool weights_dirty, const ForwardArgs& a, cudaStream_t stream, int* launches) {
  const char* why = "";
  int nl = 0;
  if (*state == nullptr) {
    *state = new TcState();
    weights_dirty = true;
  }
  TcState* st = *state;
  const int L = md.n_layers;
  if (weights_dirty) {
    for (int l = 0; l < L; ++l) {
      TcLayerImg& li = st->layers[l];
      TcLayerParams p;
      SVD_REQUIRE(tc_layer_params(md, l, p, &why), "tensor-core engine: %s", why);
      if (li.a1u) cudaFree(li.a1u);
      if (li.a1w) cudaFree(li.a1w);
      if (li.a2) cudaFree(li.a2);
      if (li.bias) cudaFree(li.bias);
      li.a1u = li.a1w = li.a2 = nullptr;
      li.bias = nullptr;
      SVD_CUDA_TRY(cudaMalloc(&li.a1u, p.a1u_bytes));
      if (p.has_s1w) SVD_CUDA_TRY(cudaMalloc(&li.a1w, p.a1w_bytes));
      SVD_CUDA_TRY(cudaMalloc(&li.a2, p.a2_bytes));
      SVD_CUDA_TRY(cudaMalloc(&li.bias, sizeof(float) * 4 * p.H));
      const LayerDesc& Ld = md.layers[l];
      const Block& bw = Ld.blocks[0];
      const Block& bu = Ld.blocks[1];
      pack_a1_kernel<<<64, 256, 0, stream>>>(bu.left, bu.left_ld, bu.scale, bu.rank, round_up(bu.rank, 8), p.H,
                                             reinterpret_cast<__nv_bfloat16*>(li.a1u));
      if (p.has_s1w)
        pack_a1_kernel<<<64, 256, 0, stream>>>(bw.left, bw.left_ld, bw.scale, bw.rank, round_up(bw.rank, 8), p.Kin,
                                               reinterpret_cast<__nv_bfloat16*>(li.a1w));
      pack_a2_kernel<<<296, 256, 0, stream>>>(bw, bu, p.H, p.ru_pad, p.rin_pad, l == 0 ? 1 : 0, Ld.d_in,
                                              reinterpret_cast<__nv_bfloat16*>(li.a2), li.bias, Ld.bias);
      nl += p.has_s1w ? 3 : 2;
      p.a1u = li.a1u;
      p.a1w = li.a1w;
      p.a2 = li.a2;
      p.bias = li.bias;
      li.prm = p;
    }
    st->n_layers = L;
    SVD_CUDA_TRY(cudaGetLastError());
  }
  const int B = a.B, T = a.T;
  const int n_cta = (B + kN - 1) / kN;
  // workspaces
  const int Dpad = st->layers[0].prm.Kin;
  const size_t xbytes = (size_t)n_cta * T * act_tile_bytes(Dpad);
  if (st->xseq_bytes < xbytes) {
    if (st->xseq) cudaFree(st->xseq);
    SVD_CUDA_TRY(cudaMalloc(&st->xseq, xbytes));
    st->xseq_bytes = xbytes;
  }
  for (int l = 0; l < L; ++l) {
    const size_t hb = (size_t)n_cta * T * act_tile_bytes(st->layers[l].prm.H);
    const int slot = l & 1;
    if (st->seq_bytes[slot] < hb) {
      if (st->seq[slot]) cudaFree(st->seq[slot]);
      SVD_CUDA_TRY(cudaMalloc(&st->seq[slot], hb));
      st->seq_bytes[slot] = hb;
    }
  }
  pack_x_kernel<<<dim3(64, n_cta), 256, 0, stream>>>(a.x, B, T, md.input_dim, Dpad, st->xseq);
  ++nl;
  static long long* dbg_buf = nullptr;
  const char* dbg_env = getenv("SVDLSTM_TC_TIMELINE");
  if (dbg_env && !dbg_buf) SVD_CUDA_TRY(cudaMalloc(&dbg_buf, sizeof(long long) * 64 * 16 * kMaxLayers));
  for (int l = 0; l < L; ++l) {
    TcLayerParams p = st->layers[l].prm;
    p.T = T;
    p.dbg = dbg_env ? dbg_buf + (size_t)l * 64 * 16 : nullptr;
    p.in_seq = (l == 0) ? st->xseq : st->seq[(l - 1) & 1];
    p.out_seq = st->seq[l & 1];
    const TcSmemPlan sp = tc_plan(p);
    SVD_CUDA_TRY(cudaFuncSetAttribute(lstm_tc_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp.total));
    lstm_tc_layer_kernel<<<n_cta, kTcThreads, sp.total, stream>>>(p);
    ++nl;
  }
  unpack_out_kernel<<<dim3(128, n_cta), 256, 0, stream>>>(st->seq[(L - 1) & 1], B, T, st->layers[L - 1].prm.H, md.dense_kernel,
                                                          md.dense_bias, md.n_out, a.y);
  ++nl;
  SVD_CUDA_TRY(cudaGetLastError());
  if (dbg_env) {   // debugging aid: dump the per-step timeline of CTA 0 (cycles relative to the step's first stamp)
    static long long host[64 * 16 * kMaxLayers];
    SVD_CUDA_TRY(cudaStreamSynchronize(stream));
    SVD_CUDA_TRY(cudaMemcpy(host, dbg_buf, sizeof(long long) * 64 * 16 * L, cudaMemcpyDeviceToHost));
    for (int l = 0; l < L; ++l)
      for (int t = 20; t < 24 && t < T; ++t) {
        const long long* r = host + ((size_t)l * 64 + t) * 16;
        fprintf(stderr, "[tc timeline] layer %d step %d:", l, t);
        for (int i = 0; i < 15; ++i) fprintf(stderr, " %lld", r[i] ? r[i] - r[0] : -1);
        fprintf(stderr, "  | step period %lld\n", r[0] - (r - 16)[0]);
      }
  }
  *launches = nl;
  return 0;
}

}  // namespace svdlstm
