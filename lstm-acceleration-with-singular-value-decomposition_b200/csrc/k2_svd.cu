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
// K2b: B = (U_r * S_r) V1 ; C = V1^-1 V2 by Gauss-Jordan with partial pivoting on [V1 | V2] (float64)
// ------------------------------------------------------------------------------------------------
__global__ void rf_init(const float* __restrict__ V, int ldv, int r, int n, double* M) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (size_t)r * n; idx += stride) {
    const int i = idx / n, c = idx - (size_t)i * n;
    M[idx] = (double)V[(size_t)i * ldv + c];
  }
}

// B[i][c] = sum_k U[i][k] S[k] V[k][c], c < r   (uses the ORIGINAL V, before elimination)
__global__ void rf_make_B(const float* __restrict__ Ur, int ldu, const float* __restrict__ Sr, const float* __restrict__ V, int ldv,
                          int m, int r, float* B) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (size_t)m * r; idx += stride) {
    const int i = idx / r, c = idx - (size_t)i * r;
    double acc = 0.0;
    for (int k = 0; k < r; ++k) acc += ((double)Ur[(size_t)i * ldu + k] * (double)Sr[k]) * (double)V[(size_t)k * ldv + c];
    B[idx] = (float)acc;
  }
}

struct RfCtl {
  int pivot_row;
  int pad;
  double pivot_val;
  double min_abs, max_abs;
};

__global__ void __launch_bounds__(256) rf_find_pivot(const double* M, int r, int n, int step, RfCtl* ctl) {
  __shared__ double sv[256];
  __shared__ int si[256];
  double best = -1.0;
  int bi = step;
  for (int i = step + threadIdx.x; i < r; i += 256) {
    const double v = fabs(M[(size_t)i * n + step]);
    if (v > best) { best = v; bi = i; }
  }
  sv[threadIdx.x] = best;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const double o = sv[threadIdx.x + s];
      const int oi = si[threadIdx.x + s];
      if (o > sv[threadIdx.x] || (o == sv[threadIdx.x] && oi < si[threadIdx.x])) { sv[threadIdx.x] = o; si[threadIdx.x] = oi; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    ctl->pivot_row = si[0];
    ctl->pivot_val = M[(size_t)si[0] * n + step];
    const double a = sv[0];
    if (step == 0) { ctl->min_abs = a; ctl->max_abs = a; }
    else { ctl->min_abs = fmin(ctl->min_abs, a); ctl->max_abs = fmax(ctl->max_abs, a); }
  }
}

// swap rows step<->pivot, scale pivot row, stash multipliers: split in two launches to avoid races
__global__ void rf_swap_scale(double* M, int r, int n, int step, const RfCtl* ctl, double* mult) {
  const int pr = ctl->pivot_row;
  const double inv = 1.0 / ctl->pivot_val;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < (size_t)n; c += stride) {
    const double a = M[(size_t)step * n + c];
    const double b = M[(size_t)pr * n + c];
    M[(size_t)pr * n + c] = a;          // (pr == step: harmless)
    M[(size_t)step * n + c] = b * inv;
  }
}
__global__ void rf_stash_mult(const double* M, int r, int n, int step, double* mult) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < r) mult[i] = (i == step) ? 0.0 : M[(size_t)i * n + step];
}
__global__ void rf_eliminate(double* M, int r, int n, int step, const double* mult) {
  const int i = blockIdx.y;
  const double f = mult[i];
  if (f == 0.0) return;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < (size_t)n; c += stride)
    M[(size_t)i * n + c] -= f * M[(size_t)step * n + c];
}
__global__ void rf_write_C(const double* M, int r, int n, float* C, const RfCtl* ctl, float* pivot_ratio) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const int nc = n - r;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (size_t)r * nc; idx += stride) {
    const int i = idx / nc, c = idx - (size_t)i * nc;
    C[idx] = (float)M[(size_t)i * n + r + c];
  }
  if (pivot_ratio && blockIdx.x == 0 && threadIdx.x == 0) *pivot_ratio = (float)(ctl->max_abs > 0 ? ctl->min_abs / ctl->max_abs : 0.0);
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

extern "C" int svdlstm_reduce_factors(const float* U_r, int ldu, const float* S_r, const float* V_r, int ldv, int m, int r, int n,
                                      float* B, float* C, float* pivot_ratio, void* stream_) {
  SVD_REQUIRE(U_r && S_r && V_r && B, "svdlstm_reduce_factors: null argument");
  SVD_REQUIRE(m >= 1 && r >= 1 && n >= r, "svdlstm_reduce_factors: need m>=1, 1<=r<=n (m=%d r=%d n=%d)", m, r, n);
  SVD_REQUIRE(C != nullptr || n == r, "svdlstm_reduce_factors: null C");
  cudaStream_t stream = (cudaStream_t)stream_;
  double *M = nullptr, *mult = nullptr;
  RfCtl* ctl = nullptr;
  SVD_CUDA_TRY(cudaMallocAsync(&M, sizeof(double) * r * n, stream));
  SVD_CUDA_TRY(cudaMallocAsync(&mult, sizeof(double) * r, stream));
  SVD_CUDA_TRY(cudaMallocAsync(&ctl, sizeof(RfCtl), stream));
  const int gx = (int)(((size_t)r * n + 255) / 256 < 1184 ? ((size_t)r * n + 255) / 256 : 1184);
  rf_init<<<gx, 256, 0, stream>>>(V_r, ldv, r, n, M);
  rf_make_B<<<(int)(((size_t)m * r + 127) / 128 < 1184 ? ((size_t)m * r + 127) / 128 : 1184), 128, 0, stream>>>(U_r, ldu, S_r, V_r, ldv, m, r, B);
  const int gc = (n + 255) / 256;
  for (int step = 0; step < r; ++step) {
    rf_find_pivot<<<1, 256, 0, stream>>>(M, r, n, step, ctl);
    rf_swap_scale<<<gc, 256, 0, stream>>>(M, r, n, step, ctl, mult);
    rf_stash_mult<<<(r + 127) / 128, 128, 0, stream>>>(M, r, n, step, mult);
    rf_eliminate<<<dim3(gc, r), 256, 0, stream>>>(M, r, n, step, mult);
  }
  if (n > r || pivot_ratio) rf_write_C<<<gx, 256, 0, stream>>>(M, r, n, C, ctl, pivot_ratio);
  SVD_CUDA_TRY(cudaGetLastError());
  SVD_CUDA_TRY(cudaFreeAsync(M, stream));
  SVD_CUDA_TRY(cudaFreeAsync(mult, stream));
  SVD_CUDA_TRY(cudaFreeAsync(ctl, stream));
  return 0;
}
