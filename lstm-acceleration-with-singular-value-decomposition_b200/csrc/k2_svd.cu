// K2: batched one-sided (Hestenes) Jacobi SVD, and K2b: reduced-factor construction B, C.
//
// Replaces np.linalg.svd(mat, full_matrices=False) (reference code/svd_classes_v3.py:491,562;
// old_versions/svd_classes.py:10,15,231) and the (U*S)@V1 / inv(V1)@V2 construction of
// make_LSTM_reduced_model (:622-626, :656-660).
//
// The k = min(m,n) short-side vectors (rows of A when m<=n, columns otherwise) are orthogonalised
// by plane rotations in a round-robin tournament ordering (k-1 rounds of k/2 disjoint pairs per
// sweep); the same rotations are accumulated in a k x k matrix J.  At convergence the vector norms
// are the singular values.  Working precision is float64 (inputs/outputs float32), so singular
// values come out at ~1e-7 relative (float32 rounding of the result) -- the 1e-5 bar of the spec
// holds for the smallest sigma too.
//   small path : one CTA per matrix, G and J in shared memory, a warp per pair;
//   large path : G, J in global (L2-resident) scratch, one launch per round, a CTA per pair.
#include <vector>

#include "common.cuh"

namespace svdlstm {
namespace {

constexpr int kSvdThreads = 256;
constexpr int kMaxSweeps = 40;
constexpr double kTol = 1e-13;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// round-robin tournament: pair `i` of round `r` among n_even players (circle method)
__device__ __forceinline__ void rr_pair(int r, int i, int n_even, int& p, int& q) {
  const int m = n_even - 1;
  if (i == 0) {
    p = r % m;
    q = m;
  } else {
    p = (r + i) % m;
    q = (r - i + m) % m;
  }
  if (p > q) {
    const int t = p;
    p = q;
    q = t;
  }
}

__device__ __forceinline__ bool jacobi_cs(double alpha, double beta, double gamma, double& c, double& s) {
  if (alpha == 0.0 || beta == 0.0) return false;
  if (fabs(gamma) <= kTol * sqrt(alpha * beta)) return false;
  const double zeta = (beta - alpha) / (2.0 * gamma);
  const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  c = 1.0 / sqrt(1.0 + t * t);
  s = c * t;
  return true;
}

// Writes S, U, Vt for one matrix from converged G (k x len), J (k x k).  Called by a whole CTA.
__device__ void svd_finalize(const double* G, const double* J, double* sig /*k, scratch*/, int* rnk /*k, scratch*/,
                             int m, int n, int k, int len, bool wide, float* U, float* S, float* Vt) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  // norms + sign (largest |entry| of the long vector made positive)
  for (int i = warp; i < k; i += nwarps) {
    double a = 0.0, best = 0.0;
    int besti = 0x7fffffff;
    for (int c = lane; c < len; c += 32) {
      const double v = G[(size_t)i * len + c];
      a += v * v;
      if (fabs(v) > fabs(best) || (fabs(v) == fabs(best) && c < besti)) { best = v; besti = c; }
    }
    a = warp_sum(a);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, s);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, s);
      if (fabs(ob) > fabs(best) || (fabs(ob) == fabs(best) && oi < besti)) { best = ob; besti = oi; }
    }
    if (lane == 0) sig[i] = (best < 0.0 ? -1.0 : 1.0) * sqrt(a);  // sign carried in sig
  }
  __syncthreads();
  for (int i = tid; i < k; i += nthr) {
    const double si = fabs(sig[i]);
    int r = 0;
    for (int j = 0; j < k; ++j) {
      const double sj = fabs(sig[j]);
      r += (sj > si) || (sj == si && j < i);
    }
    rnk[i] = r;
    S[r] = (float)si;
  }
  __syncthreads();
  for (int i = 0; i < k; ++i) {
    const int r = rnk[i];
    const double si = fabs(sig[i]);
    const double sgn = sig[i] < 0.0 ? -1.0 : 1.0;
    const double inv = si > 0.0 ? sgn / si : 0.0;
    if (wide) {
      if (Vt) for (int c = tid; c < len; c += nthr) Vt[(size_t)r * n + c] = (float)(G[(size_t)i * len + c] * inv);
      if (U) for (int c = tid; c < k; c += nthr) U[(size_t)c * k + r] = (float)(J[(size_t)i * k + c] * sgn);
    } else {
      if (U) for (int c = tid; c < len; c += nthr) U[(size_t)c * k + r] = (float)(G[(size_t)i * len + c] * inv);
      if (Vt) for (int c = tid; c < k; c += nthr) Vt[(size_t)r * n + c] = (float)(J[(size_t)i * k + c] * sgn);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// small path: everything in shared memory
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSvdThreads) svd_small_kernel(const float* __restrict__ A, int m, int n, float* U, float* S,
                                                                float* Vt, int* sweeps_out) {
  extern __shared__ __align__(16) double sm[];
  const int k = min(m, n), len = max(m, n);
  const bool wide = m <= n;
  double* G = sm;
  double* J = G + (size_t)k * len;
  double* sig = J + (size_t)k * k;
  int* rnk = reinterpret_cast<int*>(sig + k);
  __shared__ int s_changed;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwarps = kSvdThreads / 32;
  const float* Ab = A + (size_t)blockIdx.x * m * n;
  for (int idx = tid; idx < k * len; idx += kSvdThreads) {
    const int i = idx / len, c = idx - i * len;
    G[idx] = wide ? (double)Ab[(size_t)i * n + c] : (double)Ab[(size_t)c * n + i];
  }
  for (int idx = tid; idx < k * k; idx += kSvdThreads) J[idx] = (idx / k == idx % k) ? 1.0 : 0.0;
  __syncthreads();

  const int n_even = k + (k & 1);
  const int npairs = n_even / 2;
  int sweep = 0;
  for (; sweep < kMaxSweeps; ++sweep) {
    if (tid == 0) s_changed = 0;
    __syncthreads();
    for (int r = 0; r < n_even - 1; ++r) {
      for (int pi = warp; pi < npairs; pi += nwarps) {
        int p, q;
        rr_pair(r, pi, n_even, p, q);
        if (q >= k) continue;
        double* gp = G + (size_t)p * len;
        double* gq = G + (size_t)q * len;
        double al = 0.0, be = 0.0, ga = 0.0;
        for (int c = lane; c < len; c += 32) {
          const double x = gp[c], y = gq[c];
          al += x * x;
          be += y * y;
          ga += x * y;
        }
        al = warp_sum(al);
        be = warp_sum(be);
        ga = warp_sum(ga);
        double cs, sn;
        if (jacobi_cs(al, be, ga, cs, sn)) {
          for (int c = lane; c < len; c += 32) {
            const double x = gp[c], y = gq[c];
            gp[c] = cs * x - sn * y;
            gq[c] = sn * x + cs * y;
          }
          double* jp = J + (size_t)p * k;
          double* jq = J + (size_t)q * k;
          for (int c = lane; c < k; c += 32) {
            const double x = jp[c], y = jq[c];
            jp[c] = cs * x - sn * y;
            jq[c] = sn * x + cs * y;
          }
          if (lane == 0) s_changed = 1;
        }
      }
      __syncthreads();
    }
    const int ch = s_changed;
    __syncthreads();
    if (!ch) break;
  }
  // sweeps_out > 0: sweeps used (the last one found nothing to rotate); < 0: the cap was hit without converging
  if (tid == 0 && sweeps_out) sweeps_out[blockIdx.x] = sweep < kMaxSweeps ? sweep + 1 : -kMaxSweeps;
  const size_t b = blockIdx.x;
  svd_finalize(G, J, sig, rnk, m, n, k, len, wide, U ? U + b * m * k : nullptr, S + b * k, Vt ? Vt + b * k * n : nullptr);
}

// ------------------------------------------------------------------------------------------------
// large path: global scratch, one launch per round
// ------------------------------------------------------------------------------------------------
struct LargeCtl {  // one per matrix
  int changed;
  int done;
  int sweeps;
  int pad;
};

__global__ void svd_large_init(const float* __restrict__ A, int m, int n, double* G, double* J, LargeCtl* ctl) {
  const int k = min(m, n), len = max(m, n);
  const bool wide = m <= n;
  const size_t b = blockIdx.y;
  const float* Ab = A + b * m * n;
  double* Gb = G + b * k * len;
  double* Jb = J + b * k * k;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (size_t)k * len; idx += stride) {
    const int i = idx / len, c = idx - (size_t)i * len;
    Gb[idx] = wide ? (double)Ab[(size_t)i * n + c] : (double)Ab[(size_t)c * n + i];
  }
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (size_t)k * k; idx += stride)
    Jb[idx] = (idx / k == idx % k) ? 1.0 : 0.0;
  if (blockIdx.x == 0 && threadIdx.x == 0) ctl[b] = LargeCtl{0, 0, 0, 0};
}

// One Jacobi rotation of the row pair `pair` of round `round` of matrix b (whole CTA; every thread takes the same early exits).
__device__ __forceinline__ void svd_large_rotate(double* G, double* J, LargeCtl* ctl, int k, int len, int round, int pair, size_t b) {
  if (ctl[b].done) return;
  const int n_even = k + (k & 1);
  int p, q;
  rr_pair(round, pair, n_even, p, q);
  if (q >= k) return;
  double* gp = G + b * k * len + (size_t)p * len;
  double* gq = G + b * k * len + (size_t)q * len;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double al = 0.0, be = 0.0, ga = 0.0;
  for (int c = tid; c < len; c += kSvdThreads) {
    const double x = gp[c], y = gq[c];
    al += x * x;
    be += y * y;
    ga += x * y;
  }
  __shared__ double red[3][kSvdThreads / 32];
  __shared__ double s_cs[2];
  __shared__ int s_rot;
  al = warp_sum(al);
  be = warp_sum(be);
  ga = warp_sum(ga);
  if (lane == 0) {
    red[0][warp] = al;
    red[1][warp] = be;
    red[2][warp] = ga;
  }
  __syncthreads();
  if (tid == 0) {
    double a2 = 0, b2 = 0, g2 = 0;
    for (int w = 0; w < kSvdThreads / 32; ++w) {
      a2 += red[0][w];
      b2 += red[1][w];
      g2 += red[2][w];
    }
    double cs = 1.0, sn = 0.0;
    const bool rot = jacobi_cs(a2, b2, g2, cs, sn);
    s_cs[0] = cs;
    s_cs[1] = sn;
    s_rot = rot;
    if (rot) ctl[b].changed = 1;
  }
  __syncthreads();
  if (!s_rot) return;
  const double cs = s_cs[0], sn = s_cs[1];
  for (int c = tid; c < len; c += kSvdThreads) {
    const double x = gp[c], y = gq[c];
    gp[c] = cs * x - sn * y;
    gq[c] = sn * x + cs * y;
  }
  double* jp = J + b * k * k + (size_t)p * k;
  double* jq = J + b * k * k + (size_t)q * k;
  for (int c = tid; c < k; c += kSvdThreads) {
    const double x = jp[c], y = jq[c];
    jp[c] = cs * x - sn * y;
    jq[c] = sn * x + cs * y;
  }
}

// fallback when a cooperative launch is not available: one launch per round
__global__ void __launch_bounds__(kSvdThreads) svd_large_round(double* G, double* J, LargeCtl* ctl, int k, int len, int round) {
  svd_large_rotate(G, J, ctl, k, len, round, (int)blockIdx.x, (size_t)blockIdx.y);
}

// Grid-wide barrier of a cooperative (fully co-resident) launch: arrival counter + generation, sense-reversing.
__device__ __forceinline__ void svd_grid_barrier(unsigned int* bar, unsigned int n_cta) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    volatile unsigned int* gen_p = bar + 1;
    const unsigned int gen = *gen_p;
    if (atomicAdd(bar, 1u) == n_cta - 1u) {
      bar[0] = 0u;
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      while (*gen_p == gen) {
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// ALL sweeps of ALL matrices in ONE cooperative launch (the per-round version needs (k-1) x 16 launches per call: 16 368 for the
// 1024 x 4096 factors of an H = 1024 layer, most of them no-ops after convergence).  CTAs stride over the (pair, matrix) items of a
// round, meet at a grid barrier between rounds (pairs of one round are disjoint, rounds are not), and leave as soon as every
// matrix has converged.
__global__ void __launch_bounds__(kSvdThreads) svd_large_persistent(double* G, double* J, LargeCtl* ctl, int k, int len, int batch, int max_sweeps,
                                                                   unsigned int* gbar) {
  const int n_even = k + (k & 1), pairs = n_even / 2, items = pairs * batch;
  __shared__ int s_all_done;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    for (int round = 0; round < n_even - 1; ++round) {
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        svd_large_rotate(G, J, ctl, k, len, round, item % pairs, (size_t)(item / pairs));
        __syncthreads();   // the rotation's shared scratch is reused by the next item
      }
      svd_grid_barrier(gbar, gridDim.x);
    }
    if (blockIdx.x == 0)
      for (int b = threadIdx.x; b < batch; b += kSvdThreads)
        if (!ctl[b].done) {
          ctl[b].sweeps += 1;
          if (!ctl[b].changed) ctl[b].done = 1;
          ctl[b].changed = 0;
        }
    svd_grid_barrier(gbar, gridDim.x);
    if (threadIdx.x == 0) {
      int all = 1;
      for (int b = 0; b < batch; ++b) all &= ((volatile LargeCtl*)ctl)[b].done;
      s_all_done = all;
    }
    __syncthreads();
    if (s_all_done) break;
  }
}

__global__ void svd_large_sweep_end(LargeCtl* ctl, int batch) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  if (ctl[b].done) return;
  ctl[b].sweeps += 1;
  if (!ctl[b].changed) ctl[b].done = 1;
  ctl[b].changed = 0;
}

__global__ void __launch_bounds__(kSvdThreads) svd_large_finalize(const double* G, const double* J, double* sigbuf, int* rnkbuf,
                                                                  const LargeCtl* ctl, int m, int n, float* U, float* S, float* Vt,
                                                                  int* sweeps_out) {
  const int k = min(m, n), len = max(m, n);
  const size_t b = blockIdx.x;
  if (threadIdx.x == 0 && sweeps_out) sweeps_out[b] = ctl[b].done ? ctl[b].sweeps : -ctl[b].sweeps;   // < 0: not converged
  svd_finalize(G + b * k * len, J + b * k * k, sigbuf + b * k, rnkbuf + b * k, m, n, k, len, m <= n, U ? U + b * m * k : nullptr,
               S + b * k, Vt ? Vt + b * k * n : nullptr);
}

// ------------------------------------------------------------------------------------------------
// K2b: B = (U_r * S_r) V1 ; C = V1^-1 V2  for a BATCH of matrices in ONE launch (float64 inside).
//   CTA kinds (blockIdx.x order: all inverse CTAs first, so that a tile CTA that waits for an inverse can never keep
//   that inverse's CTA from being scheduled):
//     inverse  one CTA per item: in-place Gauss-Jordan inversion of V1 (r x r) with partial pivoting, in shared memory
//              when r <= 158, else in the item's L2-resident scratch; publishes inv + the pivot ratio, then a flag;
//     C tile   64 columns of C = inv . V2 (waits for the item's flag);
//     B tile   64 rows of B = (U_r * S_r) . V1 (independent of the inverse).
//   Round 1 issued 2 + 4 r launches per matrix (514 for one rank-128 factor; ~5e5 for a 2-factor rank sweep).
// ------------------------------------------------------------------------------------------------
constexpr int kRfThreads = 256;
constexpr int kRfTile = 64;
constexpr int kRfSmemRank = 158;   // r*r doubles + bookkeeping <= 200 KB

struct RfItem {   // device copy of svdlstm_reduce_item + scratch
  const float* U;
  const float* S;
  const float* V;
  float* B;
  float* C;
  float* pivot_ratio;
  double* inv;        // r x r
  unsigned int* flag;
  int ldu, ldv, m, r, n;
};
struct RfTile {
  int item;
  int kind;   // 1: C tile (first column c0), 2: B tile (first row c0)
  int c0;
};

__device__ void rf_invert(const RfItem& it, double* A /* r x r, shared or global */, int* perm /* r, shared */) {
  const int r = it.r, tid = threadIdx.x;
  __shared__ double s_val[kRfThreads / 32];
  __shared__ int s_idx[kRfThreads / 32];
  __shared__ double s_piv, s_min, s_max;
  __shared__ int s_prow;
  for (int idx = tid; idx < r * r; idx += kRfThreads) {
    const int i = idx / r, c = idx - i * r;
    A[idx] = (double)it.V[(size_t)i * it.ldv + c];
  }
  __syncthreads();
  for (int k = 0; k < r; ++k) {
    // ---- pivot: largest |A[i][k]|, i >= k (ties: smallest row) ----
    double best = -1.0;
    int bi = k;
    for (int i = k + tid; i < r; i += kRfThreads) {
      const double v = fabs(A[(size_t)i * r + k]);
      if (v > best) { best = v; bi = i; }
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, sft);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, sft);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if ((tid & 31) == 0) { s_val[tid >> 5] = best; s_idx[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kRfThreads / 32; ++w)
        if (s_val[w] > best || (s_val[w] == best && s_idx[w] < bi)) { best = s_val[w]; bi = s_idx[w]; }
      s_prow = bi;
      s_piv = A[(size_t)bi * r + k];
      perm[k] = bi;
      if (k == 0) { s_min = best; s_max = best; }
      else { s_min = fmin(s_min, best); s_max = fmax(s_max, best); }
    }
    __syncthreads();
    const int pr = s_prow;
    const double inv_p = 1.0 / s_piv;
    // ---- swap rows k <-> pr, scale the pivot row; its column-k entry becomes 1/pivot (the in-place inverse) ----
    for (int c = tid; c < r; c += kRfThreads) {
      const double a = A[(size_t)k * r + c], b2 = A[(size_t)pr * r + c];
      A[(size_t)pr * r + c] = a;                                  // (pr == k: harmless)
      A[(size_t)k * r + c] = (c == k ? 1.0 : b2) * inv_p;
    }
    __syncthreads();
    // ---- eliminate column k from every other row ----
    for (int i = tid >> 5; i < r; i += kRfThreads / 32) {
      if (i == k) continue;
      const double f = A[(size_t)i * r + k];
      __syncwarp();
      if (f != 0.0)
        for (int c = tid & 31; c < r; c += 32) {
          const double pk = A[(size_t)k * r + c];
          A[(size_t)i * r + c] = (c == k ? 0.0 : A[(size_t)i * r + c]) - f * pk;
        }
      __syncwarp();
    }
    __syncthreads();
  }
  // ---- undo the row interchanges as column interchanges, last first ----
  for (int k = r - 1; k >= 0; --k) {
    const int pr = perm[k];
    if (pr != k)
      for (int i = tid; i < r; i += kRfThreads) {
        const double a = A[(size_t)i * r + k];
        A[(size_t)i * r + k] = A[(size_t)i * r + pr];
        A[(size_t)i * r + pr] = a;
      }
    __syncthreads();
  }
  if (A != it.inv)
    for (int idx = tid; idx < r * r; idx += kRfThreads) it.inv[idx] = A[idx];
  if (tid == 0 && it.pivot_ratio) *it.pivot_ratio = (float)(s_max > 0 ? s_min / s_max : 0.0);
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    atomicExch(it.flag, 1u);
  }
}

__global__ void __launch_bounds__(kRfThreads) rf_batched_kernel(const RfItem* __restrict__ items, int n_items, const RfTile* __restrict__ tiles) {
  extern __shared__ __align__(16) unsigned char rf_smem[];
  const int tid = threadIdx.x;
  if ((int)blockIdx.x < n_items) {
    const RfItem it = items[blockIdx.x];
    double* A = it.r <= kRfSmemRank ? reinterpret_cast<double*>(rf_smem) : it.inv;
    int* perm = reinterpret_cast<int*>(rf_smem + (it.r <= kRfSmemRank ? sizeof(double) * it.r * it.r : 0));
    rf_invert(it, A, perm);
    return;
  }
  const RfTile tl = tiles[blockIdx.x - n_items];
  const RfItem it = items[tl.item];
  const int r = it.r;
  float* tile = reinterpret_cast<float*>(rf_smem);
  if (tl.kind == 2) {
    // B rows [c0, c0+64): stage (U * S) rows, then thread = output column
    const int rows = min(kRfTile, it.m - tl.c0);
    for (int idx = tid; idx < rows * r; idx += kRfThreads) {
      const int i = idx / r, k = idx - i * r;
      tile[idx] = it.U[(size_t)(tl.c0 + i) * it.ldu + k] * it.S[k];
    }
    __syncthreads();
    for (int c = tid; c < r; c += kRfThreads)
      for (int i = 0; i < rows; ++i) {
        double acc = 0.0;
        for (int k = 0; k < r; ++k) acc += (double)tile[i * r + k] * (double)it.V[(size_t)k * it.ldv + c];
        it.B[(size_t)(tl.c0 + i) * r + c] = (float)acc;
      }
    return;
  }
  // C columns [c0, c0+64) of the n - r columns of V2
  const int nc = it.n - r, cols = min(kRfTile, nc - tl.c0);
  for (int idx = tid; idx < r * kRfTile; idx += kRfThreads) {
    const int k = idx / kRfTile, c = idx - k * kRfTile;
    tile[idx] = c < cols ? it.V[(size_t)k * it.ldv + r + tl.c0 + c] : 0.f;
  }
  if (tid == 0) {
    const long long t0 = clock64();
    while (atomicAdd(it.flag, 0u) == 0u)
      if (clock64() - t0 > 20000000000LL) __trap();   // a protocol bug must surface as an error, never as a hang
    __threadfence();
  }
  __syncthreads();
  const int c = tid % kRfTile, g = tid / kRfTile;     // 4 row groups
  if (c < cols)
    for (int i = g; i < r; i += kRfThreads / kRfTile) {
      const double* row = it.inv + (size_t)i * r;
      double acc = 0.0;
      for (int k = 0; k < r; ++k) acc += __ldcg(row + k) * (double)tile[k * kRfTile + c];
      it.C[(size_t)i * nc + tl.c0 + c] = (float)acc;
    }
}

}  // namespace
}  // namespace svdlstm

using namespace svdlstm;

extern "C" int svdlstm_svd_jacobi_batched(const float* A, int batch, int m, int n, float* U, float* S, float* Vt, int* sweeps,
                                          void* stream_) {
  SVD_REQUIRE(A && S, "svdlstm_svd_jacobi_batched: null A / S");
  SVD_REQUIRE(batch >= 1 && m >= 1 && n >= 1, "svdlstm_svd_jacobi_batched: batch=%d m=%d n=%d", batch, m, n);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int k = m < n ? m : n, len = m < n ? n : m;
  const size_t small_bytes = ((size_t)k * len + (size_t)k * k + k) * sizeof(double) + (size_t)k * sizeof(int) + 16;
  if (small_bytes <= 200 * 1024) {
    SVD_CUDA_TRY(cudaFuncSetAttribute(svd_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_bytes));
    svd_small_kernel<<<batch, kSvdThreads, small_bytes, stream>>>(A, m, n, U, S, Vt, sweeps);
    SVD_CUDA_TRY(cudaGetLastError());
    return 0;
  }
  // ONE stream-ordered scratch block (freed on every exit path by the guard below)
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t oG = 0, oJ = oG + up(sizeof(double) * batch * k * len), oS = oJ + up(sizeof(double) * batch * k * k),
               oR = oS + up(sizeof(double) * batch * k), oC = oR + up(sizeof(int) * batch * k), oB = oC + up(sizeof(LargeCtl) * batch),
               total = oB + 256;
  uint8_t* scratch = nullptr;
  SVD_CUDA_TRY(cudaMallocAsync(&scratch, total, stream));
  struct Guard {
    uint8_t* p;
    cudaStream_t s;
    ~Guard() { cudaFreeAsync(p, s); }
  } guard{scratch, stream};
  double* G = reinterpret_cast<double*>(scratch + oG);
  double* J = reinterpret_cast<double*>(scratch + oJ);
  double* sig = reinterpret_cast<double*>(scratch + oS);
  int* rnk = reinterpret_cast<int*>(scratch + oR);
  LargeCtl* ctl = reinterpret_cast<LargeCtl*>(scratch + oC);
  unsigned int* gbar = reinterpret_cast<unsigned int*>(scratch + oB);
  svd_large_init<<<dim3(296, batch), 256, 0, stream>>>(A, m, n, G, J, ctl);
  const int n_even = k + (k & 1);
  // Same cap as the small path.  The persistent kernel leaves as soon as every matrix has converged (a cyclic Jacobi on
  // float32-exact data needs ~8-12 sweeps), so the cap costs nothing; a matrix that hits it is reported through `sweeps` (< 0).
  const int max_sweeps = kMaxSweeps;
  bool persistent = false;
  {
    int dev = 0, coop = 0, n_sm = 0, per_sm = 0;
    SVD_CUDA_TRY(cudaGetDevice(&dev));
    SVD_CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    SVD_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    SVD_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, svd_large_persistent, kSvdThreads, 0));
    if (coop && per_sm > 0 && !getenv("SVDLSTM_SVD_PER_ROUND")) {
      SVD_CUDA_TRY(cudaMemsetAsync(gbar, 0, 2 * sizeof(unsigned int), stream));
      const int items = (n_even / 2) * batch;
      int grid = n_sm * per_sm;
      if (grid > items) grid = items;
      int k_ = k, len_ = len, batch_ = batch, ms_ = max_sweeps;
      void* args[] = {&G, &J, &ctl, &k_, &len_, &batch_, &ms_, &gbar};
      SVD_CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(svd_large_persistent), dim3((unsigned)grid), dim3(kSvdThreads), args, 0, stream));
      persistent = true;
    }
  }
  if (!persistent)
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
      for (int r = 0; r < n_even - 1; ++r) svd_large_round<<<dim3(n_even / 2, batch), kSvdThreads, 0, stream>>>(G, J, ctl, k, len, r);
      svd_large_sweep_end<<<(batch + 127) / 128, 128, 0, stream>>>(ctl, batch);
    }
  svd_large_finalize<<<batch, kSvdThreads, 0, stream>>>(G, J, sig, rnk, ctl, m, n, U, S, Vt, sweeps);
  SVD_CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int svdlstm_reduce_factors_batched(const svdlstm_reduce_item* items, int n_items, void* stream_) {
  SVD_REQUIRE(items != nullptr && n_items >= 1, "svdlstm_reduce_factors_batched: null / empty item list");
  cudaStream_t stream = (cudaStream_t)stream_;
  std::vector<RfItem> hit((size_t)n_items);
  std::vector<RfTile> tiles;
  size_t inv_doubles = 0;
  int max_r = 0;
  for (int i = 0; i < n_items; ++i) {
    const svdlstm_reduce_item& s = items[i];
    SVD_REQUIRE(s.U_r && s.S_r && s.V_r && s.B, "svdlstm_reduce_factors: null argument (item %d)", i);
    SVD_REQUIRE(s.m >= 1 && s.r >= 1 && s.n >= s.r, "svdlstm_reduce_factors: need m>=1, 1<=r<=n (m=%d r=%d n=%d, item %d)", s.m, s.r, s.n, i);
    SVD_REQUIRE(s.C != nullptr || s.n == s.r, "svdlstm_reduce_factors: null C (item %d)", i);
    SVD_REQUIRE(s.r <= 1024, "svdlstm_reduce_factors: rank %d above 1024 (item %d)", s.r, i);
    hit[i] = RfItem{s.U_r, s.S_r, s.V_r, s.B, s.C, s.pivot_ratio, nullptr, nullptr, s.ldu, s.ldv, s.m, s.r, s.n};
    inv_doubles += (size_t)s.r * s.r;
    max_r = s.r > max_r ? s.r : max_r;
    for (int c0 = 0; c0 < s.n - s.r; c0 += kRfTile) tiles.push_back(RfTile{i, 1, c0});
    for (int r0 = 0; r0 < s.m; r0 += kRfTile) tiles.push_back(RfTile{i, 2, r0});
  }
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t o_inv = 0, o_flag = o_inv + up(sizeof(double) * inv_doubles), o_items = o_flag + up(sizeof(unsigned) * n_items),
               o_tiles = o_items + up(sizeof(RfItem) * n_items), total = o_tiles + up(sizeof(RfTile) * (tiles.size() + 1));
  uint8_t* scratch = nullptr;
  SVD_CUDA_TRY(cudaMallocAsync(&scratch, total, stream));
  struct Guard {
    uint8_t* p;
    cudaStream_t s;
    ~Guard() { cudaFreeAsync(p, s); }
  } guard{scratch, stream};
  size_t off = 0;
  for (int i = 0; i < n_items; ++i) {
    hit[i].inv = reinterpret_cast<double*>(scratch + o_inv) + off;
    hit[i].flag = reinterpret_cast<unsigned*>(scratch + o_flag) + i;
    off += (size_t)hit[i].r * hit[i].r;
  }
  SVD_CUDA_TRY(cudaMemsetAsync(scratch + o_flag, 0, sizeof(unsigned) * n_items, stream));
  // pageable sources: the runtime stages them before the call returns
  SVD_CUDA_TRY(cudaMemcpyAsync(scratch + o_items, hit.data(), sizeof(RfItem) * n_items, cudaMemcpyHostToDevice, stream));
  if (!tiles.empty()) SVD_CUDA_TRY(cudaMemcpyAsync(scratch + o_tiles, tiles.data(), sizeof(RfTile) * tiles.size(), cudaMemcpyHostToDevice, stream));
  size_t smem = sizeof(float) * (size_t)kRfTile * max_r;                                          // tile CTAs
  const int rs = max_r <= kRfSmemRank ? max_r : kRfSmemRank;
  const size_t inv_smem = sizeof(double) * (size_t)rs * rs + sizeof(int) * (size_t)max_r + 16;    // inverse CTAs
  if (inv_smem > smem) smem = inv_smem;
  SVD_REQUIRE(smem <= 227 * 1024, "svdlstm_reduce_factors: rank %d needs %zu bytes of shared memory", max_r, smem);
  SVD_CUDA_TRY(cudaFuncSetAttribute(rf_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rf_batched_kernel<<<(unsigned)(n_items + tiles.size()), kRfThreads, smem, stream>>>(reinterpret_cast<const RfItem*>(scratch + o_items), n_items,
                                                                                       reinterpret_cast<const RfTile*>(scratch + o_tiles));
  SVD_CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int svdlstm_reduce_factors(const float* U_r, int ldu, const float* S_r, const float* V_r, int ldv, int m, int r, int n,
                                      float* B, float* C, float* pivot_ratio, void* stream_) {
  svdlstm_reduce_item it{U_r, S_r, V_r, B, C, pivot_ratio, ldu, ldv, m, r, n};
  return svdlstm_reduce_factors_batched(&it, 1, stream_);
}
