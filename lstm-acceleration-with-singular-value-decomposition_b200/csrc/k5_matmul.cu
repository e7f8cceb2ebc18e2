// K5: small scaled matrix product  out = (A * diag(scale)) . B + bias  (FP32, 64 x 64 register-blocked tiles).
//
// The two dense products of the path that are NOT inside the recurrent kernels:
//   - rank truncation of a weight matrix, A_r = (U * s_r) V  (reference code/old_versions/svd_classes.py:9-12, 210-217:
//     `np.matmul(u * s, v)` after zeroing the trailing singular values) -- once per greedy-sweep iteration;
//   - a stand-alone Dense / TimeDistributed(Dense) layer called outside a fused model (svd_classes_v3.py:532-539).
// Round 1 used library GEMMs for both; these are a few MFLOP each, so a plain shared-memory tiled kernel is the right size.
#include "common.cuh"

namespace svdlstm {
namespace {

constexpr int kMT = 64, kNT = 64, kKT = 16;

__global__ void __launch_bounds__(256) scaled_matmul_kernel(const float* __restrict__ A, int lda, const float* __restrict__ scale,
                                                            const float* __restrict__ Bm, int ldb, const float* __restrict__ bias, int m, int k,
                                                            int n, float* __restrict__ out, int ldo) {
  __shared__ float As[kKT][kMT + 4];
  __shared__ float Bs[kKT][kNT + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;   // 16 x 16 threads, 4 x 4 outputs each
  const int i0 = blockIdx.y * kMT, c0 = blockIdx.x * kNT;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < k; k0 += kKT) {
    for (int idx = tid; idx < kMT * kKT; idx += 256) {
      const int i = idx / kKT, kk = idx - i * kKT;
      const int gi = i0 + i, gk = k0 + kk;
      As[kk][i] = (gi < m && gk < k) ? A[(size_t)gi * lda + gk] * (scale ? scale[gk] : 1.f) : 0.f;
    }
    for (int idx = tid; idx < kKT * kNT; idx += 256) {
      const int kk = idx / kNT, c = idx - kk * kNT;
      const int gk = k0 + kk, gc = c0 + c;
      Bs[kk][c] = (gk < k && gc < n) ? Bm[(size_t)gk * ldb + gc] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kKT; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = As[kk][ty * 4 + u];
        b[u] = Bs[kk][tx * 4 + u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int gi = i0 + ty * 4 + u, gc = c0 + tx * 4 + v;
      if (gi < m && gc < n) out[(size_t)gi * ldo + gc] = acc[u][v] + (bias ? bias[gc] : 0.f);
    }
}

}  // namespace
}  // namespace svdlstm

using namespace svdlstm;

extern "C" int svdlstm_scaled_matmul(const float* A, int lda, const float* scale, const float* B, int ldb, const float* bias, int m, int k, int n,
                                     float* out, int ldo, void* stream_) {
  SVD_REQUIRE(A && B && out, "svdlstm_scaled_matmul: null argument");
  SVD_REQUIRE(m >= 1 && n >= 1 && k >= 0 && lda >= k && ldb >= n && ldo >= n, "svdlstm_scaled_matmul: bad shape (m=%d k=%d n=%d lda=%d ldb=%d ldo=%d)", m, k, n,
              lda, ldb, ldo);
  scaled_matmul_kernel<<<dim3((n + kNT - 1) / kNT, (m + kMT - 1) / kMT), 256, 0, (cudaStream_t)stream_>>>(A, lda, scale, B, ldb, bias, m, k, n, out, ldo);
  SVD_CUDA_TRY(cudaGetLastError());
  return 0;
}
