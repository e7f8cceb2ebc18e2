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
#include <vector>

#include "common.cuh"

namespace svdlstm {
namespace {

constexpr int kChunk = 4096;
constexpr int kTile = 32;
constexpr int kSplitF = 256;   // features per Gram CTA
constexpr int kPartDoubles = kTile * kTile + 2 * kTile;   // one parked partial: the tile + both row-norm blocks

struct Work {
  int item, kind, a, b;
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
  if (wk.kind == 0) {
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
      const int kind = work[gw].kind;
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

extern "C" int svdlstm_penalties(const svdlstm_penalty_item* items, int n_items, double* out, void* stream_) {
  SVD_REQUIRE(items && out, "svdlstm_penalties: null argument");
  SVD_REQUIRE(n_items >= 1 && n_items <= 4096, "svdlstm_penalties: n_items=%d not in [1,4096]", n_items);
  cudaStream_t stream = (cudaStream_t)stream_;
  std::vector<ItemDev> hi(n_items);
  std::vector<Work> hw;
  std::vector<int> soff;          // per split tile: offset of its scratch block (doubles)
  size_t scratch_doubles = 0;
  for (int i = 0; i < n_items; ++i) {
    const svdlstm_penalty_item& s = items[i];
    SVD_REQUIRE(s.data && s.rows >= 1 && s.cols >= 1 && s.ld >= s.cols, "svdlstm_penalties: item %d has bad shape (%d,%d) ld=%d", i, s.rows, s.cols, s.ld);
    hi[i] = ItemDev{s.data, s.rows, s.cols, s.ld, s.gram, s.columns, (int)hw.size(), 0};
    const size_t total = (size_t)s.rows * s.cols;
    const int nchunks = (int)((total + kChunk - 1) / kChunk);
    for (int c = 0; c < nchunks; ++c) hw.push_back(Work{i, 0, c, 0, 0, 1, 0, 0});
    if (s.gram) {
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
  const int n_work = (int)hw.size();
  ItemDev* di = nullptr;
  Work* dw = nullptr;
  double* partial = nullptr;
  unsigned int* ticket = nullptr;
  double* scratch = nullptr;
  int* dsoff = nullptr;
  const size_t n_tick = 1 + soff.size();
  SVD_CUDA_TRY(cudaMallocAsync(&di, sizeof(ItemDev) * n_items, stream));
  SVD_CUDA_TRY(cudaMallocAsync(&dw, sizeof(Work) * n_work, stream));
  SVD_CUDA_TRY(cudaMallocAsync(&partial, sizeof(double) * 2 * n_work, stream));
  SVD_CUDA_TRY(cudaMallocAsync(&ticket, sizeof(unsigned int) * n_tick, stream));
  SVD_CUDA_TRY(cudaMallocAsync(&scratch, sizeof(double) * (scratch_doubles ? scratch_doubles : 1), stream));
  SVD_CUDA_TRY(cudaMallocAsync(&dsoff, sizeof(int) * (soff.size() ? soff.size() : 1), stream));
  SVD_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned int) * n_tick, stream));
  SVD_CUDA_TRY(cudaMemsetAsync(partial, 0, sizeof(double) * 2 * n_work, stream));
  if (!soff.empty()) SVD_CUDA_TRY(cudaMemcpyAsync(dsoff, soff.data(), sizeof(int) * soff.size(), cudaMemcpyHostToDevice, stream));
  // pageable -> device: the runtime stages these synchronously, so the vectors may die after the call
  SVD_CUDA_TRY(cudaMemcpyAsync(di, hi.data(), sizeof(ItemDev) * n_items, cudaMemcpyHostToDevice, stream));
  SVD_CUDA_TRY(cudaMemcpyAsync(dw, hw.data(), sizeof(Work) * n_work, cudaMemcpyHostToDevice, stream));
  penalties_kernel<<<n_work, 256, 0, stream>>>(di, n_items, dw, n_work, partial, ticket, scratch, dsoff, out);
  SVD_CUDA_TRY(cudaGetLastError());
  SVD_CUDA_TRY(cudaFreeAsync(di, stream));
  SVD_CUDA_TRY(cudaFreeAsync(dw, stream));
  SVD_CUDA_TRY(cudaFreeAsync(partial, stream));
  SVD_CUDA_TRY(cudaFreeAsync(ticket, stream));
  SVD_CUDA_TRY(cudaFreeAsync(scratch, stream));
  SVD_CUDA_TRY(cudaFreeAsync(dsoff, stream));
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
