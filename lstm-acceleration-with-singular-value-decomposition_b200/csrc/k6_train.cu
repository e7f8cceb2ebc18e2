// K6: training step of the 3-factor model -- forward with cache, back-propagation through time, regulariser gradients, Adam.
//
// Replaces `smodel.fit(...)` of the reference driver (code/svd_acceleration_v3.py:119-128: Hoyer fine-tune of the singular
// values with loss="mse", optimizer="adam") for models of SingularLSTMCells (svd_classes_v3.py:17-236; merged or split) +
// Dense top.  Trainable weights follow the reference: kernel (sigma_w) and recurrent_kernel (sigma_u) always (:40,:47), the
// four factor matrices and the bias only with train_uv (:49-56, :102-112), the Dense top always (stock Keras layer).
//
// One CTA owns one sequence of the mini-batch for the whole step: forward over t = 0..T-1 writing (q, gates, c, h) of every
// layer-step to an HBM cache, then backward over t = T-1..0.  Every cell form is the canonical block list of common.cuh
// (p = scale * (in . left); z += p . right), so for a block
//     d scale[k]    += q[k] dp[k]                    q = in . left (cached), dp = dz . right^T
//     d right[k][n] += scale[k] q[k] dz[n]           d left[i][k] += in[i] scale[k] dp[k]
//     d in[i]        = sum_k left[i][k] scale[k] dp[k]
// Gradients are accumulated WITHOUT atomics into the CTA's own slice of a (batch x parameters) buffer and summed over the
// batch in fixed order by reduce_grads_kernel: bit-reproducible.  FP32 throughout (the tensor-core engine is inference only).
#include <math.h>

#include <vector>

#include "common.cuh"

namespace svdlstm {
namespace {

constexpr int kTrThreads = 256;
constexpr int kTrWarps = kTrThreads / 32;

struct TrBlock {
  const float* left;
  const float* scale;
  const float* right;
  int left_ld, right_ld, rank, ncols, out0, from_h, p_off;
  int g_left_ld, g_right_ld;            // row strides of the weight tensors = of their gradients (left_ld / right_ld change when
                                        // the kernel stages the block compactly in shared memory)
  long long g_left, g_scale, g_right;   // offsets into the flat gradient vector, -1 = not trainable
};
struct TrLayer {
  int d_in, units, n_blocks, p_total;
  const float* bias;
  long long g_bias;
  long long cache_off;    // floats, per (sequence, step): start of this layer's record [q | i f g o | c | h]
  TrBlock blocks[kMaxBlocks];
};
struct TrModel {
  int n_layers, input_dim, n_out;
  const float* dense_kernel;
  const float* dense_bias;
  long long g_dense_kernel, g_dense_bias;
  long long n_params;
  long long cache_stride;   // floats per (sequence, step)
  int max_units, max_p, max_in;
  int stage_floats;         // > 0: every weight of the model fits shared memory (this many floats) and is staged there once per launch
  int grad_floats;          // > 0 (only with stage_floats): the CTA's gradient slice accumulates in shared memory too, written out once
  TrLayer layers[kMaxLayers];
};

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// grid = batch; dynamic smem: see train_smem_floats()
__global__ void __launch_bounds__(kTrThreads) train_step_kernel(const TrModel* __restrict__ mp, const float* __restrict__ x,
                                                                const float* __restrict__ y_true, int B, int T, int ret_seq,
                                                                float* __restrict__ cache, float* __restrict__ ypred,
                                                                float* __restrict__ gpart, float* __restrict__ loss_part) {
  extern __shared__ __align__(16) float sm[];
  __shared__ TrModel M;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  for (int i = tid; i < (int)(sizeof(TrModel) / 4); i += kTrThreads) reinterpret_cast<uint32_t*>(&M)[i] = reinterpret_cast<const uint32_t*>(mp)[i];
  __syncthreads();
  const int L = M.n_layers, HM = M.max_units;
  // smem map
  float* hst = sm;                       // [L][HM]  h_l(t-1) / h_l(t)
  float* cst = hst + L * HM;             // [L][HM]
  float* dhn = cst + L * HM;             // [L][HM]  d h_l from step t+1
  float* dcn = dhn + L * HM;             // [L][HM]
  float* q = dcn + L * HM;               // [max_p]
  float* dp = q + M.max_p;               // [max_p]
  float* z = dp + M.max_p;               // [4 HM]   pre-activations / dz
  float* dha = z + 4 * HM;               // [max(HM, max_in)] gradient arriving from the layer above
  float* dhb = dha + (HM > M.max_in ? HM : M.max_in);   // [same] gradient for the layer below (double buffer)
  float* vin = dhb + (HM > M.max_in ? HM : M.max_in);   // [max_in] staged input vector
  float* part = vin + M.max_in;          // [kMaxBlocks][max(HM, max_in)] per-block partial input gradients (narrow layers)
  int* kblk = reinterpret_cast<int*>(part + kMaxBlocks * (HM > M.max_in ? HM : M.max_in));   // [L][max_p] block owning projection k
  float* dy = reinterpret_cast<float*>(kblk + L * M.max_p);   // [n_out or HM]
  const int n_y = M.n_out > 0 ? M.n_out : M.layers[L - 1].units;
  const int n_steps_out = ret_seq ? T : 1;
  const float inv_count = 1.0f / ((float)B * (float)n_steps_out * (float)n_y);
  float* cb = cache + (size_t)b * T * M.cache_stride;
  float* gp = gpart + (size_t)b * M.n_params;
  float* gp_out = nullptr;
  float* yp = ypred + (size_t)b * n_steps_out * n_y;

  for (int i = tid; i < 4 * L * HM; i += kTrThreads) hst[i] = 0.f;   // hst, cst, dhn, dcn
  for (int i = tid; i < L * M.max_p; i += kTrThreads) {
    const TrLayer& Ly = M.layers[i / M.max_p];
    const int k = i % M.max_p;
    int bi = 0;
    while (bi + 1 < Ly.n_blocks && k >= Ly.blocks[bi + 1].p_off) ++bi;
    kblk[i] = bi;
  }
  // Small models (the reference's 3 x 15 fine-tune): every factor, bias and the Dense top are copied ONCE into shared memory,
  // compactly per block, and the block table (itself in shared memory) is re-pointed at the copies: the ~10 dependent phases
  // of a layer-step then wait on shared-memory instead of L2 latencies.  The gradient strides keep the tensors' own layout.
  if (M.stage_floats > 0) {
    float* wbuf = dy + (M.n_out > 0 ? M.n_out : HM) + 8;
    if (M.grad_floats > 0 && gpart != nullptr) {
      // per-step "+=" on a gradient entry is a dependent read-modify-write: in shared memory it costs ~30 cycles, in L2 ~700
      gp_out = gp;
      gp = wbuf + M.stage_floats;
      for (int i = tid; i < M.grad_floats; i += kTrThreads) gp[i] = 0.f;
    }
    __syncthreads();
    int off = 0;
    for (int l = 0; l < L; ++l) {
      TrLayer& Ly = M.layers[l];
      for (int bi = 0; bi < Ly.n_blocks; ++bi) {
        TrBlock& bk = Ly.blocks[bi];
        const int n_in = bk.from_h ? Ly.units : Ly.d_in;
        float* l_s = wbuf + off;
        float* s_s = l_s + n_in * bk.rank;
        float* r_s = s_s + bk.rank;
        for (int idx = tid; idx < n_in * bk.rank; idx += kTrThreads) l_s[idx] = bk.left[(size_t)(idx / bk.rank) * bk.left_ld + idx % bk.rank];
        for (int idx = tid; idx < bk.rank; idx += kTrThreads) s_s[idx] = bk.scale[idx];
        for (int idx = tid; idx < bk.rank * bk.ncols; idx += kTrThreads) r_s[idx] = bk.right[(size_t)(idx / bk.ncols) * bk.right_ld + idx % bk.ncols];
        off += n_in * bk.rank + bk.rank + bk.rank * bk.ncols;
        __syncthreads();
        if (tid == 0) {
          bk.left = l_s; bk.scale = s_s; bk.right = r_s;
          bk.left_ld = bk.rank; bk.right_ld = bk.ncols;
        }
      }
      float* b_s = wbuf + off;
      for (int idx = tid; idx < 4 * Ly.units; idx += kTrThreads) b_s[idx] = Ly.bias[idx];
      off += 4 * Ly.units;
      __syncthreads();
      if (tid == 0) Ly.bias = b_s;
    }
    if (M.n_out > 0) {
      const int HL = M.layers[L - 1].units;
      float* k_s = wbuf + off;
      float* d_s = k_s + HL * M.n_out;
      for (int idx = tid; idx < HL * M.n_out; idx += kTrThreads) k_s[idx] = M.dense_kernel[idx];
      for (int idx = tid; idx < M.n_out; idx += kTrThreads) d_s[idx] = M.dense_bias[idx];
      __syncthreads();
      if (tid == 0) { M.dense_kernel = k_s; M.dense_bias = d_s; }
    }
  }
  __syncthreads();

  // =============================== forward ===============================
  for (int t = 0; t < T; ++t) {
    float* ct = cb + (size_t)t * M.cache_stride;
    for (int l = 0; l < L; ++l) {
      const TrLayer& Ly = M.layers[l];
      const int H = Ly.units, D = Ly.d_in;
      float* rec = ct + Ly.cache_off;
      const float* xin = l == 0 ? x + ((size_t)b * T + t) * D : hst + (l - 1) * HM;
      for (int i = tid; i < D; i += kTrThreads) vin[i] = xin[i];
      __syncthreads();
      // q[k] = in . left[:, k]
      for (int k = tid; k < Ly.p_total; k += kTrThreads) {
        const TrBlock& bk = Ly.blocks[kblk[l * M.max_p + k]];
        const int kk = k - bk.p_off;
        const float* src = bk.from_h ? hst + l * HM : vin;
        const int n_in = bk.from_h ? H : D;
        float acc = 0.f;
        for (int i = 0; i < n_in; ++i) acc = fmaf(src[i], bk.left[(size_t)i * bk.left_ld + kk], acc);
        q[k] = acc;
        rec[k] = acc;
      }
      __syncthreads();
      for (int n = tid; n < 4 * H; n += kTrThreads) {
        float acc = Ly.bias[n];
        for (int bi = 0; bi < Ly.n_blocks; ++bi) {
          const TrBlock& bk = Ly.blocks[bi];
          const int rel = n - bk.out0;
          if (rel < 0 || rel >= bk.ncols) continue;
          for (int kk = 0; kk < bk.rank; ++kk) acc = fmaf(bk.scale[kk] * q[bk.p_off + kk], bk.right[(size_t)kk * bk.right_ld + rel], acc);
        }
        z[n] = acc;
      }
      __syncthreads();
      for (int j = tid; j < H; j += kTrThreads) {
        const float ig = 1.f / (1.f + expf(-z[j])), fg = 1.f / (1.f + expf(-z[H + j]));
        const float gg = tanhf(z[2 * H + j]), og = 1.f / (1.f + expf(-z[3 * H + j]));
        const float c = fmaf(fg, cst[l * HM + j], ig * gg);
        const float h = og * tanhf(c);
        cst[l * HM + j] = c;
        hst[l * HM + j] = h;
        float* g = rec + Ly.p_total;
        g[j] = ig; g[H + j] = fg; g[2 * H + j] = gg; g[3 * H + j] = og;
        g[4 * H + j] = c;
        g[5 * H + j] = h;
      }
      __syncthreads();
    }
    if (ret_seq || t == T - 1) {
      const int HL = M.layers[L - 1].units;
      float* yo = yp + (size_t)(ret_seq ? t : 0) * n_y;
      for (int o = tid; o < n_y; o += kTrThreads) {
        float v;
        if (M.n_out > 0) {
          v = M.dense_bias[o];
          for (int j = 0; j < HL; ++j) v = fmaf(hst[(L - 1) * HM + j], M.dense_kernel[(size_t)j * M.n_out + o], v);
        } else {
          v = hst[(L - 1) * HM + o];
        }
        yo[o] = v;
      }
      __syncthreads();
    }
  }

  // =============================== loss ==================================
  {
    float acc = 0.f;
    const float* yt = y_true + (size_t)b * n_steps_out * n_y;
    for (int i = tid; i < n_steps_out * n_y; i += kTrThreads) {
      const float d = yp[i] - yt[i];
      acc = fmaf(d, d, acc);
    }
    acc = warp_sum_f(acc);
    __shared__ float red[kTrWarps];
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int w = 0; w < kTrWarps; ++w) s += red[w];
      loss_part[b] = s * inv_count;
    }
  }
  if (gpart == nullptr) return;   // evaluation only

  // =============================== backward ==============================
  for (int i = tid; i < 2 * L * HM; i += kTrThreads) dhn[i] = 0.f;   // dhn, dcn
  __syncthreads();
  for (int t = T - 1; t >= 0; --t) {
    const float* ct = cb + (size_t)t * M.cache_stride;
    const float* cprev_t = t > 0 ? cb + (size_t)(t - 1) * M.cache_stride : nullptr;
    const int HL = M.layers[L - 1].units;
    for (int j = tid; j < HL; j += kTrThreads) dha[j] = 0.f;
    __syncthreads();
    if (ret_seq || t == T - 1) {
      const float* yo = yp + (size_t)(ret_seq ? t : 0) * n_y;
      const float* yt = y_true + ((size_t)b * n_steps_out + (ret_seq ? t : 0)) * n_y;
      for (int o = tid; o < n_y; o += kTrThreads) dy[o] = 2.f * (yo[o] - yt[o]) * inv_count;
      __syncthreads();
      const float* hL = ct + M.layers[L - 1].cache_off + M.layers[L - 1].p_total + 5 * HL;
      if (M.n_out > 0) {
        for (int idx = tid; idx < HL * M.n_out; idx += kTrThreads) {
          const int j = idx / M.n_out, o = idx - j * M.n_out;
          gp[M.g_dense_kernel + idx] += hL[j] * dy[o];
        }
        for (int o = tid; o < M.n_out; o += kTrThreads) gp[M.g_dense_bias + o] += dy[o];
        for (int j = tid; j < HL; j += kTrThreads) {
          float v = 0.f;
          for (int o = 0; o < M.n_out; ++o) v = fmaf(M.dense_kernel[(size_t)j * M.n_out + o], dy[o], v);
          dha[j] = v;
        }
      } else {
        for (int j = tid; j < HL; j += kTrThreads) dha[j] = dy[j];
      }
      __syncthreads();
    }
    float* d_cur = dha;    // gradient w.r.t. the output h_l(t) coming from above
    float* d_nxt = dhb;    // gradient for the layer below, produced by this layer
    for (int l = L - 1; l >= 0; --l) {
      const TrLayer& Ly = M.layers[l];
      const int H = Ly.units, D = Ly.d_in;
      const float* rec = ct + Ly.cache_off;
      const float* g = rec + Ly.p_total;
      const float* prev = cprev_t ? cprev_t + Ly.cache_off + Ly.p_total : nullptr;
      // dz
      for (int j = tid; j < H; j += kTrThreads) {
        const float ig = g[j], fg = g[H + j], gg = g[2 * H + j], og = g[3 * H + j], c = g[4 * H + j];
        const float c_prev = prev ? prev[4 * H + j] : 0.f;
        const float dh = d_cur[j] + dhn[l * HM + j];
        const float tc = tanhf(c);
        const float dc = fmaf(dh * og, 1.f - tc * tc, dcn[l * HM + j]);
        z[j] = dc * gg * ig * (1.f - ig);
        z[H + j] = dc * c_prev * fg * (1.f - fg);
        z[2 * H + j] = dc * ig * (1.f - gg * gg);
        z[3 * H + j] = dh * tc * og * (1.f - og);
        dcn[l * HM + j] = dc * fg;
      }
      for (int k = tid; k < Ly.p_total; k += kTrThreads) q[k] = rec[k];
      // the layer's input vector at this step (for d left)
      {
        const float* xin = l == 0 ? x + ((size_t)b * T + t) * D : ct + M.layers[l - 1].cache_off + M.layers[l - 1].p_total + 5 * M.layers[l - 1].units;
        for (int i = tid; i < D; i += kTrThreads) vin[i] = xin[i];
      }
      __syncthreads();
      // Narrow layers (the reference's 15 units): a warp per projection leaves 8 warps walking 120 projections of 15 columns in
      // turn, a shuffle reduction each -- one THREAD per projection (and per (block, input) below) finishes the same work in one
      // pass of short serial loops over shared memory.
      const bool narrow = 4 * H <= 128;
      if (narrow) {
        for (int k = tid; k < Ly.p_total; k += kTrThreads) {
          const TrBlock& bk = Ly.blocks[kblk[l * M.max_p + k]];
          const int kk = k - bk.p_off;
          const float* row = bk.right + (size_t)kk * bk.right_ld;
          const float* dz = z + bk.out0;
          float acc = 0.f;
          for (int n = 0; n < bk.ncols; ++n) acc = fmaf(dz[n], row[n], acc);
          dp[k] = acc;
          if (bk.g_scale >= 0) gp[bk.g_scale + kk] += q[k] * acc;
          if (bk.g_right >= 0) {
            const float sq = bk.scale[kk] * q[k];
            float* grow = gp + bk.g_right + (size_t)kk * bk.g_right_ld;
            for (int n = 0; n < bk.ncols; ++n) grow[n] = fmaf(sq, dz[n], grow[n]);
          }
        }
      }
      // dp[k] = sum_n dz[out0 + n] right[k][n]   (warp per k: rows of `right` are contiguous)
      for (int k = narrow ? Ly.p_total : warp; k < Ly.p_total; k += kTrWarps) {
        const TrBlock& bk = Ly.blocks[kblk[l * M.max_p + k]];
        const int kk = k - bk.p_off;
        const float* row = bk.right + (size_t)kk * bk.right_ld;
        float acc = 0.f;
        for (int n = lane; n < bk.ncols; n += 32) acc = fmaf(z[bk.out0 + n], row[n], acc);
        acc = warp_sum_f(acc);
        if (lane == 0) {
          dp[k] = acc;
          if (bk.g_scale >= 0) gp[bk.g_scale + kk] += q[k] * acc;
        }
        if (bk.g_right >= 0) {
          const float sq = bk.scale[kk] * q[k];
          float* grow = gp + bk.g_right + (size_t)kk * bk.g_right_ld;
          for (int n = lane; n < bk.ncols; n += 32) grow[n] += sq * z[bk.out0 + n];
        }
      }
      if (Ly.g_bias >= 0)
        for (int n = tid; n < 4 * H; n += kTrThreads) gp[Ly.g_bias + n] += z[n];
      __syncthreads();
      // input gradients: d in[i] = sum_k left[i][k] scale[k] dp[k]   (warp per i: rows of `left` are contiguous in k)
      const float* hprev = prev ? prev + 5 * H : nullptr;
      if (narrow) {
        const int NIN = H > D ? H : D;
        for (int e = tid; e < Ly.n_blocks * NIN; e += kTrThreads) {
          const int bi = e / NIN, i = e - bi * NIN;
          const TrBlock& bk = Ly.blocks[bi];
          const int n_in = bk.from_h ? H : D;
          if (i >= n_in || (!bk.from_h && l == 0 && bk.g_left < 0)) continue;
          const float* row = bk.left + (size_t)i * bk.left_ld;
          const float* dpk = dp + bk.p_off;
          float acc = 0.f;
          if (bk.g_left >= 0) {
            const float in_i = bk.from_h ? (hprev ? hprev[i] : 0.f) : vin[i];
            float* grow = gp + bk.g_left + (size_t)i * bk.g_left_ld;
            for (int kk = 0; kk < bk.rank; ++kk) {
              const float sdp = bk.scale[kk] * dpk[kk];
              acc = fmaf(row[kk], sdp, acc);
              grow[kk] = fmaf(in_i, sdp, grow[kk]);
            }
          } else {
            for (int kk = 0; kk < bk.rank; ++kk) acc = fmaf(row[kk], bk.scale[kk] * dpk[kk], acc);
          }
          part[e] = acc;
        }
        __syncthreads();
        for (int e = tid; e < D + H; e += kTrThreads) {
          const bool rec_side = e >= D;
          const int i = rec_side ? e - D : e;
          if (!rec_side && l == 0) continue;   // nobody needs d x
          float acc = 0.f;
          for (int bi = 0; bi < Ly.n_blocks; ++bi)
            if ((Ly.blocks[bi].from_h != 0) == rec_side) acc += part[bi * NIN + i];
          (rec_side ? dhn + l * HM : d_nxt)[i] = acc;
        }
        __syncthreads();
      }
      for (int bi = narrow ? Ly.n_blocks : 0; bi < Ly.n_blocks; ++bi) {
        const TrBlock& bk = Ly.blocks[bi];
        const int n_in = bk.from_h ? H : D;
        if (!bk.from_h && l == 0 && bk.g_left < 0) continue;   // nobody needs d x
        for (int i = warp; i < n_in; i += kTrWarps) {
          const float* row = bk.left + (size_t)i * bk.left_ld;
          const float in_i = bk.from_h ? (hprev ? hprev[i] : 0.f) : vin[i];
          float acc = 0.f;
          for (int kk = lane; kk < bk.rank; kk += 32) {
            const float sdp = bk.scale[kk] * dp[bk.p_off + kk];
            acc = fmaf(row[kk], sdp, acc);
            if (bk.g_left >= 0) gp[bk.g_left + (size_t)i * bk.g_left_ld + kk] += in_i * sdp;
          }
          acc = warp_sum_f(acc);
          if (lane == 0) {
            // several blocks (split form) feed the same input vector: the first one of each kind overwrites, the others add
            float* dst = bk.from_h ? dhn + l * HM : d_nxt;
            const bool first = bk.from_h ? (bi == Ly.n_blocks / 2) : (bi == 0);
            dst[i] = first ? acc : dst[i] + acc;
          }
        }
        __syncthreads();
      }
      float* tmp = d_cur;
      d_cur = d_nxt;
      d_nxt = tmp;
    }
  }
  if (gp_out != nullptr) {
    __syncthreads();
    for (int i = tid; i < M.grad_floats; i += kTrThreads) gp_out[i] = gp[i];
  }
}

// grad[p] = sum_b gpart[b][p] (fixed order); loss = sum_b loss_part[b]
__global__ void reduce_grads_kernel(const float* __restrict__ gpart, long long n_params, int B, float* __restrict__ grad,
                                    const float* __restrict__ loss_part, float* __restrict__ loss) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_params; p += stride) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += gpart[(size_t)b * n_params + p];
    grad[p] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += loss_part[b];
    *loss = s;
  }
}

// Regulariser terms, one CTA per item, added to grad (and to *loss):
//   kind 1  Hoyer (svd_classes_v3.py:455-462):  pen = c * S1 / S2,   d/dx_k = c (sign(x_k) / S2 - 2 x_k S1 / S2^2)
//   kind 2  Keras OrthogonalRegularizer(mode='rows') (:514,:573): pen = c * 0.5 * sum_{i != j} |xh_i . xh_j| / (n(n-1)/2)
//           with xh = x / max(||x||, 1e-6);  d/dxh_i = c * sum_{j != i} sign(P_ij) xh_j / pairs, projected through the normalisation
struct RegItem {
  const float* w;
  float* g;
  float* scratch;   // kind 2: rows x cols floats
  int rows, cols, kind;
  float coef;
};
__global__ void __launch_bounds__(256) reg_grads_kernel(const RegItem* __restrict__ items, float* __restrict__ loss) {
  const RegItem it = items[blockIdx.x];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  __shared__ float red[2][8];
  __shared__ float s_tot[2];
  const int n = it.rows * it.cols;
  if (it.kind == 1) {
    float s1 = 0.f, s2 = 0.f;
    for (int i = tid; i < n; i += 256) {
      const float v = it.w[i];
      s1 += fabsf(v);
      s2 = fmaf(v, v, s2);
    }
    s1 = warp_sum_f(s1);
    s2 = warp_sum_f(s2);
    if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
    __syncthreads();
    if (tid == 0) {
      float a = 0.f, c = 0.f;
      for (int w = 0; w < 8; ++w) { a += red[0][w]; c += red[1][w]; }
      s_tot[0] = a;
      s_tot[1] = c;
      atomicAdd(loss, it.coef * a / c);
    }
    __syncthreads();
    const float S1 = s_tot[0], S2 = s_tot[1];
    for (int i = tid; i < n; i += 256) {
      const float v = it.w[i];
      const float sg = v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f);
      it.g[i] += it.coef * (sg / S2 - 2.f * v * S1 / (S2 * S2));
    }
    return;
  }
  // kind 2: rows mode, warp per row i; g_hat_i = sum_{j != i} sign(P_ij) xh_j is built in the item's scratch row (global memory)
  const float pairs = it.rows * (it.rows - 1.0f) / 2.0f;
  if (pairs <= 0.f) return;
  float pen = 0.f;
  for (int i = warp; i < it.rows; i += 8) {
    const float* xi = it.w + (size_t)i * it.cols;
    float* gh = it.scratch + (size_t)i * it.cols;
    float ni = 0.f;
    for (int c = lane; c < it.cols; c += 32) {
      ni = fmaf(xi[c], xi[c], ni);
      gh[c] = 0.f;
    }
    ni = fmaxf(sqrtf(warp_sum_f(ni)), 1e-6f);
    for (int j = 0; j < it.rows; ++j) {
      if (j == i) continue;
      const float* xj = it.w + (size_t)j * it.cols;
      float nj = 0.f, pij = 0.f;
      for (int c = lane; c < it.cols; c += 32) {
        nj = fmaf(xj[c], xj[c], nj);
        pij = fmaf(xi[c], xj[c], pij);
      }
      nj = fmaxf(sqrtf(warp_sum_f(nj)), 1e-6f);
      pij = warp_sum_f(pij) / (ni * nj);
      pen += fabsf(pij);
      const float sg = (pij > 0.f ? 1.f : (pij < 0.f ? -1.f : 0.f)) / nj;
      for (int c = lane; c < it.cols; c += 32) gh[c] = fmaf(sg, xj[c], gh[c]);   // (each lane re-reads only what it wrote)
    }
    float dot = 0.f;
    for (int c = lane; c < it.cols; c += 32) dot = fmaf(gh[c], xi[c] / ni, dot);
    dot = warp_sum_f(dot);
    // through the normalisation xh = x / ||x||:  dx = (g_hat - xh (xh . g_hat)) / ||x||
    for (int c = lane; c < it.cols; c += 32) it.g[(size_t)i * it.cols + c] += it.coef / pairs * (gh[c] - (xi[c] / ni) * dot) / ni;
  }
  if (lane == 0) red[0][warp] = pen;     // every lane of a warp carries the same sum
  __syncthreads();
  if (tid == 0) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += red[0][w];
    atomicAdd(loss, it.coef * 0.5f * a / pairs);
  }
}

// Adam (Keras defaults unless told otherwise), all parameter tensors of the model in one launch.
struct AdamItem {
  float* w;
  const float* g;
  float* m;
  float* v;
  long long n;
};
__global__ void adam_kernel(const AdamItem* __restrict__ items, float lr_t, float b1, float b2, float eps) {
  const AdamItem it = items[blockIdx.y];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < it.n; i += stride) {
    const float g = it.g[i];
    const float m = fmaf(b1, it.m[i], (1.f - b1) * g);
    const float v = fmaf(b2, it.v[i], (1.f - b2) * g * g);
    it.m[i] = m;
    it.v[i] = v;
    it.w[i] -= lr_t * m / (sqrtf(v) + eps);
  }
}

}  // namespace
}  // namespace svdlstm

using namespace svdlstm;

struct svdlstm_trainer_s {
  svdlstm_model_s* h;
  TrModel tm;             // host copy
  TrModel* dev_tm;
  unsigned long long md_version;
  bool built;
  int train_uv[kMaxLayers];
  // flat vectors (device): grad, adam m / v
  float* grad;
  float* adam_m;
  float* adam_v;
  float* loss;            // device scalar
  // per-call scratch, grown on demand
  float* cache;
  size_t cache_floats;
  float* gpart;
  size_t gpart_floats;
  float* ypred;
  size_t ypred_floats;
  float* loss_part;
  size_t loss_part_floats;
  AdamItem* dev_adam;
  int n_adam;
  RegItem* dev_reg;
  int n_reg;
  long long step;
};

namespace {

int grow(float** p, size_t* have, size_t need) {
  if (*have >= need) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *have = 0;
  SVD_CUDA_TRY(cudaMalloc(p, sizeof(float) * need));
  *have = need;
  return 0;
}

// (Re)derive the training view of the model from the handle's block lists.  Flat parameter layout per layer:
// [sigma_w | sigma_u | w_left | w_right | u_left | u_right | bias], then [dense kernel | dense bias].
size_t train_smem_floats(const TrModel& M);

int build_train_model(svdlstm_trainer_s* tr) {
  const ModelDesc& md = tr->h->md;
  TrModel& M = tr->tm;
  memset(&M, 0, sizeof(M));
  M.n_layers = md.n_layers;
  M.input_dim = md.input_dim;
  M.n_out = md.n_out;
  M.dense_kernel = md.dense_kernel;
  M.dense_bias = md.dense_bias;
  long long off = 0, coff = 0, stage = 0;
  for (int l = 0; l < md.n_layers; ++l) {
    const LayerDesc& Ld = md.layers[l];
    TrLayer& T = M.layers[l];
    stage += 4 * Ld.units;
    SVD_REQUIRE(Ld.n_blocks == 2 || Ld.n_blocks == 8, "svdlstm_trainer: layer %d is not a factored cell", l);
    const int ng = Ld.n_blocks / 2;
    for (int bi = 0; bi < Ld.n_blocks; ++bi)
      SVD_REQUIRE(Ld.blocks[bi].left && Ld.blocks[bi].scale && !Ld.blocks[bi].ident,
                  "svdlstm_trainer: only SingularLSTMCell (3-factor) layers are trainable (layer %d)", l);
    T.d_in = Ld.d_in;
    T.units = Ld.units;
    T.n_blocks = Ld.n_blocks;
    T.p_total = Ld.p_total;
    T.bias = Ld.bias;
    const int H = Ld.units, D = Ld.d_in;
    const int kw_tot = Ld.blocks[0].left_ld, ku_tot = Ld.blocks[ng].left_ld;   // widths of w_left / u_left = total sigma counts
    const int kw = Ld.blocks[0].rank, ku = Ld.blocks[ng].rank;
    const long long o_sw = off, o_su = o_sw + kw_tot, o_wl = o_su + ku_tot, o_wr = o_wl + (long long)D * kw_tot,
                    o_ul = o_wr + (long long)kw * 4 * H, o_ur = o_ul + (long long)H * ku_tot, o_b = o_ur + (long long)ku * 4 * H;
    off = o_b + 4 * H;
    const bool uv = tr->train_uv[l] != 0;
    T.g_bias = uv ? o_b : -1;
    for (int bi = 0; bi < Ld.n_blocks; ++bi) {
      const Block& s = Ld.blocks[bi];
      TrBlock& d = T.blocks[bi];
      const bool isu = bi >= ng;
      const Block& base = Ld.blocks[isu ? ng : 0];
      d.left = s.left; d.scale = s.scale; d.right = s.right;
      d.left_ld = s.left_ld; d.right_ld = s.right_ld; d.rank = s.rank; d.ncols = s.ncols; d.out0 = s.out0; d.from_h = s.from_h; d.p_off = s.p_off;
      d.g_left_ld = s.left_ld; d.g_right_ld = s.right_ld;
      stage += (long long)(s.from_h ? H : D) * s.rank + s.rank + (long long)s.rank * s.ncols;
      d.g_scale = (isu ? o_su : o_sw) + (s.scale - base.scale);
      d.g_left = uv ? (isu ? o_ul : o_wl) + (s.left - base.left) : -1;
      d.g_right = uv ? (isu ? o_ur : o_wr) + (s.right - base.right) : -1;
    }
    T.cache_off = coff;
    coff += Ld.p_total + 6 * H;
    M.max_units = H > M.max_units ? H : M.max_units;
    M.max_p = Ld.p_total > M.max_p ? Ld.p_total : M.max_p;
    M.max_in = D > M.max_in ? D : M.max_in;
  }
  if (md.n_out > 0) {
    const int HL = md.layers[md.n_layers - 1].units;
    M.g_dense_kernel = off;
    M.g_dense_bias = off + (long long)HL * md.n_out;
    off = M.g_dense_bias + md.n_out;
  } else {
    M.g_dense_kernel = M.g_dense_bias = -1;
  }
  M.n_params = off;
  M.cache_stride = coff;
  if (md.n_out > 0) stage += (long long)md.layers[md.n_layers - 1].units * md.n_out + md.n_out;
  M.stage_floats = 0;
  M.grad_floats = 0;
  if (sizeof(float) * (train_smem_floats(M) + (size_t)stage + 16) <= 180 * 1024) {
    M.stage_floats = (int)stage + 16;
    if (sizeof(float) * (train_smem_floats(M) + (size_t)M.stage_floats + (size_t)M.n_params) <= 200 * 1024) M.grad_floats = (int)M.n_params;
  }
  return 0;
}

size_t train_smem_floats(const TrModel& M) {   // activations / scratch (the staged weights follow)
  const int HM = M.max_units, L = M.n_layers;
  const int mx = HM > M.max_in ? HM : M.max_in;
  const int n_y = M.n_out > 0 ? M.n_out : HM;
  return (size_t)4 * L * HM + 2 * (size_t)M.max_p + 4 * (size_t)HM + 2 * (size_t)mx + M.max_in + (size_t)kMaxBlocks * mx +
         (size_t)L * M.max_p + n_y + 8;
}

int sync_model(svdlstm_trainer_s* tr, cudaStream_t stream) {
  if (int e = upload_model_desc(tr->h, stream)) return e;      // (keeps the handle's own version counter moving)
  if (tr->built && tr->md_version == tr->h->md_version) return 0;
  const long long old_n = tr->built ? tr->tm.n_params : -1;
  if (int e = build_train_model(tr)) return e;
  if (!tr->dev_tm) SVD_CUDA_TRY(cudaMalloc(&tr->dev_tm, sizeof(TrModel)));
  SVD_CUDA_TRY(cudaMemcpyAsync(tr->dev_tm, &tr->tm, sizeof(TrModel), cudaMemcpyHostToDevice, stream));
  SVD_CUDA_TRY(cudaStreamSynchronize(stream));   // tm is a member: stable, but keep the upload ordered before any rebuild
  if (old_n != tr->tm.n_params) {
    if (tr->grad) cudaFree(tr->grad);
    if (tr->adam_m) cudaFree(tr->adam_m);
    tr->grad = tr->adam_m = tr->adam_v = nullptr;
    SVD_CUDA_TRY(cudaMalloc(&tr->grad, sizeof(float) * tr->tm.n_params));
    SVD_CUDA_TRY(cudaMalloc(&tr->adam_m, sizeof(float) * 2 * tr->tm.n_params));
    tr->adam_v = tr->adam_m + tr->tm.n_params;
    SVD_CUDA_TRY(cudaMemset(tr->adam_m, 0, sizeof(float) * 2 * tr->tm.n_params));
    tr->step = 0;
  }
  tr->built = true;
  tr->md_version = tr->h->md_version;
  return 0;
}

}  // namespace

extern "C" {

int svdlstm_trainer_create(svdlstm_handle h, const int* train_uv, svdlstm_trainer* out) {
  SVD_REQUIRE(h != nullptr && out != nullptr, "svdlstm_trainer_create: null argument");
  for (int l = 0; l < h->md.n_layers; ++l) SVD_REQUIRE(h->layer_set[l], "svdlstm_trainer_create: weights of layer %d were never set", l);
  svdlstm_trainer_s* tr = new (std::nothrow) svdlstm_trainer_s();
  SVD_REQUIRE(tr != nullptr, "svdlstm_trainer_create: out of host memory");
  memset(tr, 0, sizeof(*tr));
  tr->h = h;
  for (int l = 0; l < h->md.n_layers; ++l) tr->train_uv[l] = train_uv ? train_uv[l] : 0;
  if (int e = build_train_model(tr)) {
    delete tr;
    return e;
  }
  *out = tr;
  return 0;
}

void svdlstm_trainer_destroy(svdlstm_trainer tr) {
  if (!tr) return;
  cudaFree(tr->dev_tm);
  cudaFree(tr->grad);
  cudaFree(tr->adam_m);
  cudaFree(tr->loss);
  cudaFree(tr->cache);
  cudaFree(tr->gpart);
  cudaFree(tr->ypred);
  cudaFree(tr->loss_part);
  cudaFree(tr->dev_adam);
  cudaFree(tr->dev_reg);
  delete tr;
}

int64_t svdlstm_trainer_num_params(svdlstm_trainer tr) { return tr ? tr->tm.n_params : -1; }

/* Layout of the flat gradient vector: offsets[7 * l + {0..6}] = start of [sigma_w, sigma_u, w_left, w_right, u_left, u_right,
 * bias] of layer l, offsets[7 L], offsets[7 L + 1] = Dense kernel / bias, offsets[7 L + 2] = total. */
int svdlstm_trainer_layout(svdlstm_trainer tr, int64_t* offsets) {
  SVD_REQUIRE(tr != nullptr && offsets != nullptr, "svdlstm_trainer_layout: null argument");
  const ModelDesc& md = tr->h->md;
  long long off = 0;
  for (int l = 0; l < md.n_layers; ++l) {
    const LayerDesc& Ld = md.layers[l];
    const int ng = Ld.n_blocks / 2, H = Ld.units, D = Ld.d_in;
    const int kw_tot = Ld.blocks[0].left_ld, ku_tot = Ld.blocks[ng].left_ld, kw = Ld.blocks[0].rank, ku = Ld.blocks[ng].rank;
    const long long sizes[7] = {kw_tot, ku_tot, (long long)D * kw_tot, (long long)kw * 4 * H, (long long)H * ku_tot, (long long)ku * 4 * H, 4 * H};
    for (int i = 0; i < 7; ++i) {
      offsets[7 * l + i] = off;
      off += sizes[i];
    }
  }
  offsets[7 * md.n_layers] = off;
  if (md.n_out > 0) off += (long long)md.layers[md.n_layers - 1].units * md.n_out;
  offsets[7 * md.n_layers + 1] = off;
  if (md.n_out > 0) off += md.n_out;
  offsets[7 * md.n_layers + 2] = off;
  return 0;
}

/* Forward + back-propagation of the mean-squared error of one mini-batch.  x (B,T,D), y_true (B,T,n) with return_sequences
 * else (B,n), device pointers.  grad_out (device, num_params floats, may be NULL = loss only) receives d loss / d parameter
 * in the layout above (zeros for non-trainable entries); loss_out (device float) the data loss.  */
int svdlstm_trainer_gradients(svdlstm_trainer tr, const float* x, const float* y_true, int B, int T, int return_sequences,
                              float* grad_out, float* loss_out, void* stream_) {
  SVD_REQUIRE(tr != nullptr && x != nullptr && y_true != nullptr, "svdlstm_trainer_gradients: null argument");
  SVD_REQUIRE(B >= 1 && T >= 1, "svdlstm_trainer_gradients: B=%d T=%d must be >= 1", B, T);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = sync_model(tr, stream)) return e;
  const TrModel& M = tr->tm;
  const int n_y = M.n_out > 0 ? M.n_out : M.layers[M.n_layers - 1].units;
  const int n_so = return_sequences ? T : 1;
  if (int e = grow(&tr->cache, &tr->cache_floats, (size_t)B * T * M.cache_stride)) return e;
  if (int e = grow(&tr->ypred, &tr->ypred_floats, (size_t)B * n_so * n_y)) return e;
  if (int e = grow(&tr->loss_part, &tr->loss_part_floats, (size_t)B)) return e;
  if (!tr->loss) SVD_CUDA_TRY(cudaMalloc(&tr->loss, sizeof(float)));
  const bool want_grad = grad_out != nullptr;
  if (want_grad) {
    if (int e = grow(&tr->gpart, &tr->gpart_floats, (size_t)B * M.n_params)) return e;
    SVD_CUDA_TRY(cudaMemsetAsync(tr->gpart, 0, sizeof(float) * (size_t)B * M.n_params, stream));
  }
  const size_t smem = sizeof(float) * (train_smem_floats(M) + (size_t)M.stage_floats + (size_t)M.grad_floats);
  SVD_REQUIRE(smem <= 200 * 1024, "svdlstm_trainer: model too large for the training kernel's shared memory (%zu bytes)", smem);
  SVD_CUDA_TRY(cudaFuncSetAttribute(train_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  train_step_kernel<<<B, kTrThreads, smem, stream>>>(tr->dev_tm, x, y_true, B, T, return_sequences ? 1 : 0, tr->cache, tr->ypred,
                                                    want_grad ? tr->gpart : nullptr, tr->loss_part);
  float* lo = loss_out ? loss_out : tr->loss;
  if (want_grad) {
    reduce_grads_kernel<<<(int)((M.n_params + 255) / 256 < 1184 ? (M.n_params + 255) / 256 : 1184), 256, 0, stream>>>(tr->gpart, M.n_params, B, grad_out,
                                                                                                                    tr->loss_part, lo);
  } else {
    reduce_grads_kernel<<<1, 32, 0, stream>>>(nullptr, 0, B, nullptr, tr->loss_part, lo);
  }
  SVD_CUDA_TRY(cudaGetLastError());
  return 0;
}

/* Adds the regulariser gradients to grad (flat layout) and their penalties to *loss (device float).  items[i]: tensor index
 * (7 l + {0..5}) it applies to, kind (1 Hoyer, 2 orthogonal rows), coefficient. */
int svdlstm_trainer_regularizers(svdlstm_trainer tr, const int* tensor_index, const int* kind, const float* coef, int n, float* grad,
                                 float* loss, void* stream_) {
  SVD_REQUIRE(tr != nullptr && grad != nullptr && loss != nullptr, "svdlstm_trainer_regularizers: null argument");
  if (n <= 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = sync_model(tr, stream)) return e;
  const ModelDesc& md = tr->h->md;
  std::vector<int64_t> offs(7 * md.n_layers + 3);
  svdlstm_trainer_layout(tr, offs.data());
  std::vector<RegItem> items((size_t)n);
  for (int i = 0; i < n; ++i) {
    const int l = tensor_index[i] / 7, w = tensor_index[i] % 7;
    SVD_REQUIRE(l >= 0 && l < md.n_layers && w < 6, "svdlstm_trainer_regularizers: bad tensor index %d", tensor_index[i]);
    const LayerDesc& Ld = md.layers[l];
    const int ng = Ld.n_blocks / 2, H = Ld.units, D = Ld.d_in;
    const int kw_tot = Ld.blocks[0].left_ld, ku_tot = Ld.blocks[ng].left_ld, kw = Ld.blocks[0].rank, ku = Ld.blocks[ng].rank;
    const float* ptrs[6] = {Ld.blocks[0].scale, Ld.blocks[ng].scale, Ld.blocks[0].left, Ld.blocks[0].right, Ld.blocks[ng].left, Ld.blocks[ng].right};
    const int rows[6] = {1, 1, D, kw, H, ku}, cols[6] = {kw_tot, ku_tot, kw_tot, 4 * H, ku_tot, 4 * H};
    items[i] = RegItem{ptrs[w], grad + offs[tensor_index[i]], tr->gpart ? tr->gpart + offs[tensor_index[i]] : nullptr, rows[w], cols[w], kind[i], coef[i]};
    SVD_REQUIRE(kind[i] == 1 || (kind[i] == 2 && tr->gpart != nullptr), "svdlstm_trainer_regularizers: kind %d needs a preceding gradient call", kind[i]);
  }
  if (tr->n_reg < n) {
    if (tr->dev_reg) cudaFree(tr->dev_reg);
    tr->dev_reg = nullptr;
    SVD_CUDA_TRY(cudaMalloc(&tr->dev_reg, sizeof(RegItem) * n));
    tr->n_reg = n;
  }
  SVD_CUDA_TRY(cudaMemcpyAsync(tr->dev_reg, items.data(), sizeof(RegItem) * n, cudaMemcpyHostToDevice, stream));
  reg_grads_kernel<<<n, 256, 0, stream>>>(tr->dev_reg, loss);
  SVD_CUDA_TRY(cudaGetLastError());
  return 0;
}

/* One Adam update of every trainable tensor from `grad` (flat layout): w -= lr_t m / (sqrt(v) + eps), lr_t = lr sqrt(1-b2^t)/(1-b1^t)
 * (Keras Adam; defaults lr=1e-3, b1=0.9, b2=0.999, eps=1e-7).  Weights are updated IN PLACE in the caller's tensors; the
 * caller re-binds them (svdlstm_set_singular_weights / svdlstm_set_dense_top) before the next inference forward.   */
int svdlstm_trainer_adam(svdlstm_trainer tr, const float* grad, float lr, float beta1, float beta2, float eps, void* stream_) {
  SVD_REQUIRE(tr != nullptr && grad != nullptr, "svdlstm_trainer_adam: null argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = sync_model(tr, stream)) return e;
  const ModelDesc& md = tr->h->md;
  std::vector<int64_t> offs(7 * md.n_layers + 3);
  svdlstm_trainer_layout(tr, offs.data());
  std::vector<AdamItem> items;
  for (int l = 0; l < md.n_layers; ++l) {
    const LayerDesc& Ld = md.layers[l];
    const int ng = Ld.n_blocks / 2;
    float* ptrs[7] = {const_cast<float*>(Ld.blocks[0].scale), const_cast<float*>(Ld.blocks[ng].scale), const_cast<float*>(Ld.blocks[0].left),
                      const_cast<float*>(Ld.blocks[0].right), const_cast<float*>(Ld.blocks[ng].left), const_cast<float*>(Ld.blocks[ng].right),
                      const_cast<float*>(Ld.bias)};
    for (int w = 0; w < 7; ++w) {
      if (w >= 2 && !tr->train_uv[l]) continue;
      const long long o = offs[7 * l + w], n = offs[7 * l + w + 1 <= 7 * md.n_layers ? 7 * l + w + 1 : 7 * md.n_layers] - o;
      items.push_back(AdamItem{ptrs[w], grad + o, tr->adam_m + o, tr->adam_v + o, n});
    }
  }
  if (md.n_out > 0) {
    const long long ok = offs[7 * md.n_layers], ob = offs[7 * md.n_layers + 1], oe = offs[7 * md.n_layers + 2];
    items.push_back(AdamItem{const_cast<float*>(md.dense_kernel), grad + ok, tr->adam_m + ok, tr->adam_v + ok, ob - ok});
    items.push_back(AdamItem{const_cast<float*>(md.dense_bias), grad + ob, tr->adam_m + ob, tr->adam_v + ob, oe - ob});
  }
  const int n = (int)items.size();
  if (tr->n_adam < n) {
    if (tr->dev_adam) cudaFree(tr->dev_adam);
    tr->dev_adam = nullptr;
    SVD_CUDA_TRY(cudaMalloc(&tr->dev_adam, sizeof(AdamItem) * n));
    tr->n_adam = n;
  }
  SVD_CUDA_TRY(cudaMemcpyAsync(tr->dev_adam, items.data(), sizeof(AdamItem) * n, cudaMemcpyHostToDevice, stream));
  ++tr->step;
  const double t = (double)tr->step;
  const float lr_t = (float)(lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t)));
  adam_kernel<<<dim3(64, n), 256, 0, stream>>>(tr->dev_adam, lr_t, beta1, beta2, eps);
  SVD_CUDA_TRY(cudaGetLastError());
  tr->h->tc_dirty = true;      // weights changed in place: packed tensor-core images are stale
  return 0;
}

}  // extern "C"
