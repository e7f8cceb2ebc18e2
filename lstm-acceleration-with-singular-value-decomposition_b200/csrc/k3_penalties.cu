// K3: fused Hoyer-sparsity + orthogonality penalty evaluation, and K4: rank-sweep squared error.
//
// K3 replaces the per-weight regulariser calls Keras makes every training step:
// HoyerRegularizer.__call__ (reference code/svd_classes_v3.py:460-462: hoyer*sum|x|/sum x^2) on the
// sigma vectors and keras.regularizers.OrthogonalRegularizer(mode='rows') (call sites :514,:573) on
// w_left / w_right / u_left / u_right.  ONE launch walks a work list covering every item of a model:
//   kind 0  L1 / L2 partial sums over a 4096-element chunk                     (HBM-bound, 4 B/elt once)
//   kind 1  one 32x32 tile (bi<=bj) of the Gram matrix Y Y^T in float64, with the row norms of both
//           row blocks accumulated on the fly -> sum_{i!=j}|P_ij|/(|y_i||y_j|) and ||Y Y^T - I||_F^2.
//           Wide factors (more than kSplitF features, e.g. the 128 x 4096 right factors of H = 1024) are SPLIT-K:
//           each CTA covers kSplitF features and parks its partial tile in scratch; the last CTA of the tile
//           (per-tile ticket) adds the parts in split order -- |.| and the norms are nonlinear, so they are
//           applied only to the complete tile -- and reports into the tile's first work slot (fixed position)
// Partials go to a scratch buffer; the last CTA to finish (atomic ticket) reduces them per item in
// a FIXED order, so results are bit-reproducible run to run.  Raw sums are emitted so that both
// the reference's definitions (L1/L2^2, mean |off-diagonal| of the normalised Gram) and the
// north-star's (L1/||.||_2, ||U^T U - I||_F) are derivable on the host.
//
// K4 is the device half of the RMSE of svd_acceleration_v3.py:187-190: per-rank sum of squared
// errors, two-stage fixed-order float64 reduction (independent of GPU count => sweep determinism).
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace svdlstm {
namespace {

constexpr int kChunk = 4096;
constexpr int kTile = 32;
constexpr int kSplitF = 256;   // features per Gram CTA
constexpr int kPartDoubles = kTile * kTile + 2 * kTile;   // one parked partial: the tile + both row-norm blocks

// ---- kind 2: 128 x 128 Gram tiles on the tensor cores (tcgen05, kind::tf32, FP32 accumulation in TMEM) -----------------------
// The Gram matrix of the large factors (w_right / u_right of H = 1024: 128 x 4096, u_left: 1024 x 128, ...) is a SYRK:
// 2 n^2 m FLOP, 1.3 GFLOP for the C5 factor set -- round 1 did it in float64 on the CUDA cores (0.37 ms).  Here every 128 x 128
// tile of Y Y^T is accumulated by tcgen05.mma over K chunks of 32 features with the 3xTF32 split y = hi + lo
// (hi = tf32(y), lo = tf32(y - hi)):  hi hi^T + hi lo^T + lo hi^T  recovers FP32-level accuracy (~2^-21 per product) from
// TF32 operands.  Operands are staged by the CTA's own loads (coalesced 64-byte row segments -> K-major SWIZZLE_NONE core
// matrices, hi and lo planes), double-buffered against the MMAs; the same pass accumulates the row norms of both blocks.
// Split-K over the features, per-tile ticket, fixed-order combination as for kind 1 (FP32 parts, float64 final sums).
constexpr int kTcTile = 128;            // rows of both operand blocks (MMA M = N = 128)
constexpr int kTcKc = 32;               // features per staged chunk (4 MMAs of K = 8 per product)
constexpr int kTcSplitF = 512;          // features per tensor-core Gram CTA
constexpr int kTcPartFloats = kTcTile * kTcTile + 2 * kTcTile;       // parked partial: tile + squared norms of both row blocks
constexpr uint32_t kTcPlane = kTcTile * kTcKc * 4;                    // bytes of one operand plane (hi or lo) of one chunk: 16 KB
constexpr uint32_t kTcBuf = 4 * kTcPlane;                             // A hi, A lo, B hi, B lo
constexpr uint32_t kTcSmem = 2 * kTcBuf + 64;                         // two buffers + barriers

__device__ __forceinline__ uint32_t k3_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void k3_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ bool k3_mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void k3_mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!k3_mbar_try(bar, parity))
    if (clock64() - t0 > 4000000000LL) __trap();   // a protocol bug must surface as an error, never as a hang
}
__device__ __forceinline__ uint32_t k3_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// K-major SWIZZLE_NONE operand plane [128 rows x 32 k] of 4-byte elements: core matrix = 8 rows x 16 bytes;
//   elem(row, k) at (row / 8) * 1024 + (k / 4) * 128 + (row % 8) * 16 + (k % 4) * 4        LBO = 128 (next core along K), SBO = 1024
__device__ __forceinline__ uint64_t k3_desc(uint32_t saddr) {
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | ((128u >> 4) << 16);
  const uint32_t hi = ((1024u >> 4) & 0x3FFFu) | (1u << 14);     // bit 46: sm_100 descriptor version
  return ((uint64_t)hi << 32) | lo;
}
constexpr uint32_t k3_idesc_tf32() {
  return (1u << 4)                       // D format F32
         | (2u << 7) | (2u << 10)        // A, B format TF32
         | (0u << 15) | (0u << 16)       // both K-major
         | ((uint32_t)(kTcTile >> 3) << 17) | ((uint32_t)(kTcTile >> 4) << 24);
}
__device__ __forceinline__ void k3_umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(k3_idesc_tf32()), "r"(accumulate)
      : "memory");
}

struct Work {
  int item, kind, a, b;   // kind 0: L1/L2 chunk a;  1: CUDA-core 32 x 32 Gram tile (a, b);  2: tensor-core 128 x 128 Gram tile (a, b)
  int split, n_splits;   // this CTA's feature range [split * kSplitF, ...) of n_splits
  int slot;              // split tiles: index of the tile's ticket / scratch block
  int base;              // split tiles: work index that receives the tile's result
};

struct ItemDev {
  const float* data;
  int rows, cols, ld, gram, columns;
  int first_work, n_work;
};

__device__ __forceinline__ double block_sum_256(double v, double* red) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < 8; ++w) t += red[w];
  return t;  // valid in thread 0
}


// One 128 x 128 tile (bi = wk.a <= bj = wk.b) of Y Y^T over the feature range of this split, rows mode.  Returns (off, fro) in
// thread 0 when this CTA completes the tile (finish = true), else parks its part.
__device__ void gram_tile_tc(const ItemDev& it, const Work& wk, unsigned char* dsm, unsigned int* ticket, double* scratch, const int* scratch_off,
                             double* red, double& r0, double& r1, bool& finish, int& out_slot) {
  __shared__ uint32_t s_tmem;
  __shared__ float s_nrm[2][kTcTile];
  __shared__ bool s_tile_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = it.rows, F = it.cols;
  const int i0 = wk.a * kTcTile, j0 = wk.b * kTcTile;
  const bool diag = wk.a == wk.b;
  const int kbeg = wk.split * kTcSplitF;
  const int kend = kbeg + kTcSplitF < F ? kbeg + kTcSplitF : F;
  const int n_chunks = (kend - kbeg + kTcKc - 1) / kTcKc;
  const uint32_t sbase = k3_smem_u32(dsm);
  const uint32_t bar0 = sbase + 2 * kTcBuf;      // two mbarriers: "the MMAs that read buffer b have completed"
  if (tid == 0) {
    k3_mbar_init(bar0, 1);
    k3_mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(k3_smem_u32(&s_tmem)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&s_tmem);

  // loader geometry: warp w stages row groups {w, w + 8} (8 rows each); lane = (row in group, k quarter); 2 k halves of 16
  const int lrow = lane & 7, kq = lane >> 3;
  const bool vec_ok = (it.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(it.data) & 15) == 0);
  double nacc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};   // [operand A/B][row group of this warp]
  auto stage = [&](int chunk, uint32_t buf) {
    const int k0 = kbeg + chunk * kTcKc;
#pragma unroll
    for (int op = 0; op < 2; ++op) {
      if (op == 1 && diag) break;
      const int rbase = op == 0 ? i0 : j0;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int rg = warp + 8 * g;
        const int row = rbase + rg * 8 + lrow;
#pragma unroll
        for (int kh = 0; kh < 2; ++kh) {
          const int k = k0 + kh * 16 + kq * 4;
          float v[4] = {0.f, 0.f, 0.f, 0.f};
          if (row < R) {
            const float* src = it.data + (size_t)row * it.ld + k;
            if (vec_ok && k + 3 < kend) {
              const float4 q4 = __ldg(reinterpret_cast<const float4*>(src));
              v[0] = q4.x; v[1] = q4.y; v[2] = q4.z; v[3] = q4.w;
            } else {
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (k + u < kend) v[u] = __ldg(src + u);
            }
          }
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            hi[u] = k3_tf32(v[u]);
            lo[u] = k3_tf32(v[u] - __uint_as_float(hi[u]));
            nacc[op][g] += (double)v[u] * (double)v[u];
          }
          const uint32_t off = (uint32_t)rg * 1024u + (uint32_t)(kh * 4 + kq) * 128u + (uint32_t)lrow * 16u;
          const uint32_t ph = sbase + buf * kTcBuf + (uint32_t)(2 * op) * kTcPlane + off;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ph), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ph + kTcPlane), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
        }
      }
    }
  };
  for (int c = 0; c < n_chunks; ++c) {
    const uint32_t buf = (uint32_t)c & 1u;
    if (c >= 2) k3_mbar_wait(bar0 + 8u * buf, (uint32_t)((c >> 1) - 1) & 1u);   // the MMAs of chunk c-2 no longer read this buffer
    stage(c, buf);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core's async proxy
    __syncthreads();
    if (warp == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t a_hi = sbase + buf * kTcBuf, a_lo = a_hi + kTcPlane;
        const uint32_t b_hi = diag ? a_hi : a_hi + 2 * kTcPlane, b_lo = b_hi + kTcPlane;
#pragma unroll
        for (int ks = 0; ks < kTcKc / 8; ++ks) {     // K = 8 per MMA = two core matrices = 256 bytes along K
          const uint32_t o = (uint32_t)ks * 256u;
          k3_umma_tf32(tmem, k3_desc(a_hi + o), k3_desc(b_hi + o), (c > 0 || ks > 0) ? 1u : 0u);
          k3_umma_tf32(tmem, k3_desc(a_hi + o), k3_desc(b_lo + o), 1u);
          k3_umma_tf32(tmem, k3_desc(a_lo + o), k3_desc(b_hi + o), 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 8u * buf) : "memory");
      }
      __syncwarp();
    }
  }
  // all MMAs complete: the last commit of each buffer
  {
    const int last = n_chunks - 1;
    k3_mbar_wait(bar0 + 8u * ((uint32_t)last & 1u), (uint32_t)(last >> 1) & 1u);
    if (n_chunks > 1) k3_mbar_wait(bar0 + 8u * ((uint32_t)(last - 1) & 1u), (uint32_t)((last - 1) >> 1) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // squared row norms of this split: reduce over the 4 k quarters, lanes 0..7 hold 8 rows
#pragma unroll
  for (int op = 0; op < 2; ++op)
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      double v = nacc[op][g];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 8) s_nrm[op][(warp + 8 * g) * 8 + lane] = (float)v;
    }
  __syncthreads();
  if (diag && tid < kTcTile) s_nrm[1][tid] = s_nrm[0][tid];
  __syncthreads();
  // accumulator tile -> registers: thread = row (TMEM lane quarter = warp % 4), 64 of the 128 columns (half = warp / 4)
  const int q = warp & 3, half = warp >> 2;
  const int trow = q * 32 + lane;
  float acc[64];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64 + j * 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int u = 0; u < 16; ++u) acc[j * 16 + u] = __uint_as_float(r[u]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
  finish = true;
  if (wk.n_splits > 1) {
    float* mine = reinterpret_cast<float*>(scratch + (size_t)scratch_off[wk.slot]) + (size_t)wk.split * kTcPartFloats;
#pragma unroll
    for (int u = 0; u < 64; ++u) mine[(size_t)trow * kTcTile + half * 64 + u] = acc[u];
    if (tid < kTcTile) {
      mine[kTcTile * kTcTile + tid] = s_nrm[0][tid];
      mine[kTcTile * kTcTile + kTcTile + tid] = s_nrm[1][tid];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_tile_last = (atomicAdd(ticket + 1 + wk.slot, 1u) == (unsigned int)wk.n_splits - 1u);
    __syncthreads();
    finish = s_tile_last;
    if (finish) {
      __threadfence();
      const volatile float* parts = reinterpret_cast<const volatile float*>(scratch + (size_t)scratch_off[wk.slot]);
#pragma unroll
      for (int u = 0; u < 64; ++u) acc[u] = 0.f;
      float na = 0.f, nb = 0.f;
      for (int sp = 0; sp < wk.n_splits; ++sp) {       // fixed order: bit-reproducible
        const volatile float* pp = parts + (size_t)sp * kTcPartFloats;
#pragma unroll
        for (int u = 0; u < 64; ++u) acc[u] += pp[(size_t)trow * kTcTile + half * 64 + u];
        if (tid < kTcTile) {
          na += pp[kTcTile * kTcTile + tid];
          nb += pp[kTcTile * kTcTile + kTcTile + tid];
        }
      }
      __syncthreads();
      if (tid < kTcTile) {
        s_nrm[0][tid] = na;
        s_nrm[1][tid] = nb;
      }
      __syncthreads();
      out_slot = wk.base;
    }
  }
  if (!finish) return;
  double off = 0.0, fro = 0.0;
  const double mult = diag ? 1.0 : 2.0;
  const int gi = i0 + trow;
  const double ni = sqrt(fmax((double)s_nrm[0][trow], 1e-12));
  if (gi < R) {
#pragma unroll
    for (int u = 0; u < 64; ++u) {
      const int cj = half * 64 + u, gj = j0 + cj;
      if (gj < R) {
        const double p = (double)acc[u];
        if (gi != gj) {
          off += fabs(p) / (ni * sqrt(fmax((double)s_nrm[1][cj], 1e-12)));
          fro += p * p;
        } else {
          fro += (p - 1.0) * (p - 1.0);
        }
      }
    }
  }
  r0 = block_sum_256(off * mult, red);
  r1 = block_sum_256(fro * mult, red);
}

__global__ void __launch_bounds__(256) penalties_kernel(const ItemDev* __restrict__ items, int n_items, const Work* __restrict__ work,
                                                        int n_work, double* partial /*n_work x 2, zeroed*/, unsigned int* ticket /*[0] + per split tile, zeroed*/,
                                                        double* scratch /*per split tile: n_splits x kPartDoubles*/, const int* __restrict__ scratch_off,
                                                        double* out) {
  __shared__ double As[kTile][kTile + 1];   // converted to float64 ONCE per tile load: the inner loop is DFMA + LDS.64 only
  __shared__ double Bs[kTile][kTile + 1];
  __shared__ double red[8];
  __shared__ double nrmA[kTile], nrmB[kTile];
  __shared__ bool is_last;
  __shared__ bool tile_last;
  int out_slot = blockIdx.x;   // where this CTA's (r0, r1) go
  const int tid = threadIdx.x;
  const Work wk = work[blockIdx.x];
  const ItemDev it = items[wk.item];
  double r0 = 0.0, r1 = 0.0;
  if (wk.kind == 2) {
    extern __shared__ __align__(1024) unsigned char dsm[];
    bool finish = true;
    gram_tile_tc(it, wk, dsm, ticket, scratch, scratch_off, red, r0, r1, finish, out_slot);
    if (!finish) out_slot = -1;
  } else if (wk.kind == 0) {
    const size_t total = (size_t)it.rows * it.cols;
    const size_t beg = (size_t)wk.a * kChunk;
    const size_t end = beg + kChunk < total ? beg + kChunk : total;
    double l1 = 0.0, l2 = 0.0;
    for (size_t idx = beg + tid; idx < end; idx += 256) {
      const int r = idx / it.cols, c = idx - (size_t)r * it.cols;
      const float v = __ldg(it.data + (size_t)r * it.ld + c);
      l1 += fabsf(v);
      l2 += (double)v * v;
    }
    r0 = block_sum_256(l1, red);
    r1 = block_sum_256(l2, red);
  } else {
    // Gram tile (bi=wk.a, bj=wk.b), Y = X (rows mode) or X^T (columns mode)
    const int R = it.columns ? it.cols : it.rows;   // vectors
    const int F = it.columns ? it.rows : it.cols;   // features
    const int i0 = wk.a * kTile, j0 = wk.b * kTile;
    const int ti = tid >> 3, tj = (tid & 7) * 4;
    double acc[4] = {0, 0, 0, 0};
    double nacc = 0.0;
    const int kbeg = wk.split * kSplitF;
    const int kend = wk.n_splits == 1 ? F : (kbeg + kSplitF < F ? kbeg + kSplitF : F);
    for (int k0 = kbeg; k0 < kend; k0 += kTile) {
      for (int idx = tid; idx < kTile * kTile; idx += 256) {
        int rr, kk;
        if (it.columns) { kk = idx / kTile; rr = idx - kk * kTile; }   // coalesce along the vector index
        else { rr = idx / kTile; kk = idx - rr * kTile; }
        const int k = k0 + kk;
        float va = 0.f, vb = 0.f;
        if (k < kend) {
          const int ia = i0 + rr, ib = j0 + rr;
          if (ia < R) va = __ldg(it.columns ? it.data + (size_t)k * it.ld + ia : it.data + (size_t)ia * it.ld + k);
          if (ib < R) vb = __ldg(it.columns ? it.data + (size_t)k * it.ld + ib : it.data + (size_t)ib * it.ld + k);
        }
        As[rr][kk] = (double)va;
        Bs[rr][kk] = (double)vb;
      }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < kTile; ++kk) {
        const double av = As[ti][kk];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] += av * Bs[tj + u][kk];
      }
      if (tid < 2 * kTile) {
        const double* row = tid < kTile ? As[tid] : Bs[tid - kTile];
        for (int kk = 0; kk < kTile; ++kk) nacc += row[kk] * row[kk];
      }
      __syncthreads();
    }
    bool finish = true;
    if (wk.n_splits > 1) {
      // park the partial tile; the last CTA of the tile adds the parts in split order
      double* mine = scratch + (size_t)scratch_off[wk.slot] + (size_t)wk.split * kPartDoubles;
#pragma unroll
      for (int u = 0; u < 4; ++u) mine[tid * 4 + u] = acc[u];
      if (tid < 2 * kTile) mine[kTile * kTile + tid] = nacc;
      __threadfence();
      __syncthreads();
      if (tid == 0) tile_last = (atomicAdd(ticket + 1 + wk.slot, 1u) == (unsigned int)wk.n_splits - 1u);
      __syncthreads();
      finish = tile_last;
      if (finish) {
        __threadfence();
        const volatile double* parts = scratch + (size_t)scratch_off[wk.slot];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = 0.0;
        nacc = 0.0;
        for (int sp = 0; sp < wk.n_splits; ++sp) {
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] += parts[(size_t)sp * kPartDoubles + tid * 4 + u];
          if (tid < 2 * kTile) nacc += parts[(size_t)sp * kPartDoubles + kTile * kTile + tid];
        }
        out_slot = wk.base;
      }
    }
    if (tid < kTile) nrmA[tid] = sqrt(fmax(nacc, 1e-12));
    else if (tid < 2 * kTile) nrmB[tid - kTile] = sqrt(fmax(nacc, 1e-12));
    __syncthreads();
    double off = 0.0, fro = 0.0;
    const double mult = (wk.a == wk.b) ? 1.0 : 2.0;
    const int gi = i0 + ti;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int gj = j0 + tj + u;
      if (gi < R && gj < R) {
        const double p = acc[u];
        if (gi != gj) {
          off += fabs(p) / (nrmA[ti] * nrmB[tj + u]);
          fro += p * p;
        } else {
          fro += (p - 1.0) * (p - 1.0);
        }
      }
    }
    r0 = block_sum_256(off * mult, red);
    r1 = block_sum_256(fro * mult, red);
    if (!finish) out_slot = -1;   // a parked part reports nothing (its slot stays zero)
  }
  if (tid == 0) {
    if (out_slot >= 0) {
      partial[2 * (size_t)out_slot] = r0;
      partial[2 * (size_t)out_slot + 1] = r1;
    }
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // fixed-order final reduction: one warp per item, lane-strided then shuffle tree
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = warp; i < n_items; i += 8) {
    const ItemDev d = items[i];
    double s[4] = {0, 0, 0, 0};
    for (int w = lane; w < d.n_work; w += 32) {
      const int gw = d.first_work + w;
      const int kind = work[gw].kind == 0 ? 0 : 1;     // Gram tiles of either flavour report (off-diagonal sum, Frobenius sum)
      const volatile double* pp = partial + 2 * (size_t)gw;
      s[2 * kind] += pp[0];
      s[2 * kind + 1] += pp[1];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) s[u] += __shfl_xor_sync(0xffffffffu, s[u], sft);
    if (lane == 0)
      for (int u = 0; u < 4; ++u) out[4 * i + u] = s[u];
  }
}

// ---------------------------------------------------------------------------------------------
// K4
// ---------------------------------------------------------------------------------------------
constexpr int kSseChunk = 8192;

__global__ void __launch_bounds__(256) sse_partial_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t n,
                                                          int n_chunks, double* partial) {
  __shared__ double red[8];
  const int r = blockIdx.y;
  const int64_t beg = (int64_t)blockIdx.x * kSseChunk;
  const int64_t end = beg + kSseChunk < n ? beg + kSseChunk : n;
  const float* p = pred + (size_t)r * n;
  double acc = 0.0;
  for (int64_t i = beg + threadIdx.x; i < end; i += 256) {
    const double d = (double)__ldg(p + i) - (double)__ldg(target + i);
    acc += d * d;
  }
  const double t = block_sum_256(acc, red);
  if (threadIdx.x == 0) partial[(size_t)r * n_chunks + blockIdx.x] = t;
}

__global__ void __launch_bounds__(256) sse_final_kernel(const double* __restrict__ partial, int n_chunks, double* sse) {
  __shared__ double red[8];
  const int r = blockIdx.x;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n_chunks; i += 256) acc += partial[(size_t)r * n_chunks + i];
  const double t = block_sum_256(acc, red);
  if (threadIdx.x == 0) sse[r] = t;
}

}  // namespace
}  // namespace svdlstm

using namespace svdlstm;

// The work list of a model does not change from one training step to the next (same tensors, same shapes): it is built and
// uploaded once and kept per device; a repeated call costs two small memsets and the launch.
struct K3Plan {
  std::vector<svdlstm_penalty_item> key;
  bool no_tc = false;
  ItemDev* di = nullptr;
  Work* dw = nullptr;
  double* partial = nullptr;
  unsigned int* ticket = nullptr;
  double* scratch = nullptr;
  int* dsoff = nullptr;
  int n_work = 0;
  size_t n_tick = 0;
  bool any_tc = false;
  cudaEvent_t done = nullptr;   // the last launch that used these buffers
  void release() {
    if (done) cudaEventSynchronize(done);
    cudaFree(di); cudaFree(dw); cudaFree(partial); cudaFree(ticket); cudaFree(scratch); cudaFree(dsoff);
    di = nullptr; dw = nullptr; partial = nullptr; ticket = nullptr; scratch = nullptr; dsoff = nullptr;
    key.clear();
  }
};
static K3Plan g_k3_plan[16];
static std::mutex g_k3_mu;

extern "C" int svdlstm_penalties(const svdlstm_penalty_item* items, int n_items, double* out, void* stream_) {
  SVD_REQUIRE(items && out, "svdlstm_penalties: null argument");
  SVD_REQUIRE(n_items >= 1 && n_items <= 4096, "svdlstm_penalties: n_items=%d not in [1,4096]", n_items);
  cudaStream_t stream = (cudaStream_t)stream_;
  int dev = 0;
  SVD_CUDA_TRY(cudaGetDevice(&dev));
  SVD_REQUIRE(dev >= 0 && dev < 16, "svdlstm_penalties: device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lock(g_k3_mu);
  K3Plan& P = g_k3_plan[dev];
  const bool no_tc = getenv("SVDLSTM_K3_NO_TC") != nullptr;
  bool same = (int)P.key.size() == n_items && P.no_tc == no_tc;
  for (int i = 0; same && i < n_items; ++i)
    same = P.key[i].data == items[i].data && P.key[i].rows == items[i].rows && P.key[i].cols == items[i].cols && P.key[i].ld == items[i].ld &&
           P.key[i].gram == items[i].gram && P.key[i].columns == items[i].columns;
  if (!same) {
    P.release();
    std::vector<ItemDev> hi(n_items);
    std::vector<Work> hw;
    std::vector<int> soff;          // per split tile: offset of its scratch block (doubles)
    size_t scratch_doubles = 0;
    bool any_tc = false;
    for (int i = 0; i < n_items; ++i) {
      const svdlstm_penalty_item& s = items[i];
      SVD_REQUIRE(s.data && s.rows >= 1 && s.cols >= 1 && s.ld >= s.cols, "svdlstm_penalties: item %d has bad shape (%d,%d) ld=%d", i, s.rows, s.cols, s.ld);
      hi[i] = ItemDev{s.data, s.rows, s.cols, s.ld, s.gram, s.columns, (int)hw.size(), 0};
      const size_t total = (size_t)s.rows * s.cols;
      const int nchunks = (int)((total + kChunk - 1) / kChunk);
      for (int c = 0; c < nchunks; ++c) hw.push_back(Work{i, 0, c, 0, 0, 1, 0, 0});
      // tensor-core tiles for the GEMM-sized Gram matrices (rows mode); small items and columns mode stay on the float64 tiles
      const bool tc = s.gram && !s.columns && s.rows >= 64 && s.cols >= 64 && (size_t)s.rows * s.rows * s.cols >= (size_t)1 << 21 && !no_tc;
      if (tc) {
        any_tc = true;
        const int nb = (s.rows + kTcTile - 1) / kTcTile;
        const int nsp = (s.cols + kTcSplitF - 1) / kTcSplitF;
        for (int a = 0; a < nb; ++a)
          for (int b = a; b < nb; ++b) {
            if (nsp <= 1) {
              hw.push_back(Work{i, 2, a, b, 0, 1, 0, 0});
            } else {
              const int base = (int)hw.size(), slot = (int)soff.size();
              soff.push_back((int)scratch_doubles);
              scratch_doubles += ((size_t)nsp * kTcPartFloats + 1) / 2;
              for (int sp = 0; sp < nsp; ++sp) hw.push_back(Work{i, 2, a, b, sp, nsp, slot, base});
            }
          }
      } else if (s.gram) {
        const int R = s.columns ? s.cols : s.rows;
        const int F = s.columns ? s.rows : s.cols;
        const int nb = (R + kTile - 1) / kTile;
        const int nsp = (F + kSplitF - 1) / kSplitF;
        for (int a = 0; a < nb; ++a)
          for (int b = a; b < nb; ++b) {
            if (nsp <= 1) {
              hw.push_back(Work{i, 1, a, b, 0, 1, 0, 0});
            } else {
              const int base = (int)hw.size(), slot = (int)soff.size();
              soff.push_back((int)scratch_doubles);
              scratch_doubles += (size_t)nsp * kPartDoubles;
              for (int sp = 0; sp < nsp; ++sp) hw.push_back(Work{i, 1, a, b, sp, nsp, slot, base});
            }
          }
      }
      hi[i].n_work = (int)hw.size() - hi[i].first_work;
    }
    P.n_work = (int)hw.size();
    P.n_tick = 1 + soff.size();
    P.any_tc = any_tc;
    P.no_tc = no_tc;
    SVD_CUDA_TRY(cudaMalloc(&P.di, sizeof(ItemDev) * n_items));
    SVD_CUDA_TRY(cudaMalloc(&P.dw, sizeof(Work) * P.n_work));
    SVD_CUDA_TRY(cudaMalloc(&P.partial, sizeof(double) * 2 * P.n_work));
    SVD_CUDA_TRY(cudaMalloc(&P.ticket, sizeof(unsigned int) * P.n_tick));
    SVD_CUDA_TRY(cudaMalloc(&P.scratch, sizeof(double) * (scratch_doubles ? scratch_doubles : 1)));
    SVD_CUDA_TRY(cudaMalloc(&P.dsoff, sizeof(int) * (soff.size() ? soff.size() : 1)));
    if (!P.done) SVD_CUDA_TRY(cudaEventCreateWithFlags(&P.done, cudaEventDisableTiming));
    // blocking copies from pageable memory: the host vectors die with this scope
    if (!soff.empty()) SVD_CUDA_TRY(cudaMemcpy(P.dsoff, soff.data(), sizeof(int) * soff.size(), cudaMemcpyHostToDevice));
    SVD_CUDA_TRY(cudaMemcpy(P.di, hi.data(), sizeof(ItemDev) * n_items, cudaMemcpyHostToDevice));
    SVD_CUDA_TRY(cudaMemcpy(P.dw, hw.data(), sizeof(Work) * P.n_work, cudaMemcpyHostToDevice));
    P.key.assign(items, items + n_items);
  } else {
    SVD_CUDA_TRY(cudaStreamWaitEvent(stream, P.done, 0));   // a previous call on another stream may still own the scratch
  }
  SVD_CUDA_TRY(cudaMemsetAsync(P.ticket, 0, sizeof(unsigned int) * P.n_tick, stream));
  SVD_CUDA_TRY(cudaMemsetAsync(P.partial, 0, sizeof(double) * 2 * P.n_work, stream));
  const size_t dyn_smem = P.any_tc ? kTcSmem : 0;
  if (P.any_tc) SVD_CUDA_TRY(cudaFuncSetAttribute(penalties_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
  penalties_kernel<<<P.n_work, 256, dyn_smem, stream>>>(P.di, n_items, P.dw, P.n_work, P.partial, P.ticket, P.scratch, P.dsoff, out);
  SVD_CUDA_TRY(cudaGetLastError());
  SVD_CUDA_TRY(cudaEventRecord(P.done, stream));
  return 0;
}

extern "C" int svdlstm_sweep_sse(const float* pred, const float* target, int n_ranks, int64_t n, double* sse, void* stream_) {
  SVD_REQUIRE(pred && target && sse, "svdlstm_sweep_sse: null argument");
  SVD_REQUIRE(n_ranks >= 1 && n_ranks <= 65535 && n >= 1, "svdlstm_sweep_sse: n_ranks=%d n=%lld", n_ranks, (long long)n);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int n_chunks = (int)((n + kSseChunk - 1) / kSseChunk);
  double* partial = nullptr;
  SVD_CUDA_TRY(cudaMallocAsync(&partial, sizeof(double) * (size_t)n_ranks * n_chunks, stream));
  sse_partial_kernel<<<dim3(n_chunks, n_ranks), 256, 0, stream>>>(pred, target, n, n_chunks, partial);
  sse_final_kernel<<<n_ranks, 256, 0, stream>>>(partial, n_chunks, sse);
  SVD_CUDA_TRY(cudaGetLastError());
  SVD_CUDA_TRY(cudaFreeAsync(partial, stream));
  return 0;
}
