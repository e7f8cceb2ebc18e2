// K1 (wavefront): latency-optimised FP32 persistent kernel for the real-time small-model regime
// (BASELINE config 2: the shipped DROPBEAR model D=16, H=15, L=3, batch 1, T ~ 1e5).
//
// One CTA per sequence.  Warp l+1 owns layer l; warp 0 streams x_t into a shared-memory ring with
// cp.async; warp L+1 applies the Dense top and writes y.  The layers run as a WAVEFRONT: at tick tau
// layer l works on timestep tau-l, so the L layers (and the loader / output stages) overlap and
// the time per step is ONE layer's dependent chain, not L of them.  A single CTA barrier per tick
// hands h_l(t) to layer l+1 through a double-buffered shared-memory slot.
//
// Inside a layer warp (all factor matrices are staged ONCE into shared memory, transposed so that
// lane-owned output columns read conflict-free 16-byte rows):
//   stage 1  lane q owns intermediate column q:  p[q] = scale[q] * <v, LT[q,:]>     (v = x_t|h_{l-1}(t) or h_l(t-1))
//   stage 2  lane (j,s) owns G = 4/LU gates of unit j:  z = bias + <p[group], RT[n,:]>
//   stage 3  activations in the owning lanes, xor-shuffle gather, c/h update in lane j
// The 2-factor "identity prefix" columns are stored as one-hot rows of RT (exact in FP32), so every
// cell form (full, 3-factor, 2-factor; merged or split) runs the same two phases.
//
// Replaces SingularLSTMCell.call / ReducedLSTMCell.call + backend.rnn for batch-1 streaming
// (reference code/svd_classes_v3.py:116-236, 317-368, 405-434; svd_acceleration_v3.py:151).
#include "common.cuh"

namespace svdlstm {

namespace {

constexpr int kXRing = 32;      // x_t ring depth (steps)
constexpr int kPrefetch = 12;   // cp.async distance (ticks)
constexpr int kYRing = 32;      // y staging before a coalesced store

__host__ __device__ inline int round4(int x) { return (x + 3) & ~3; }
// row stride (floats): multiple of 4 with an odd number of 16-byte units => conflict-free LDS.128
__host__ __device__ inline int odd4(int x) {
  int u = (x + 3) / 4;
  if ((u & 1) == 0) ++u;
  return 4 * u;
}

struct WfLayer {
  int H, Din;
  int n_groups;        // 1 (merged/full) or 4 (split)
  int g_start[4];      // start of each group's p range (multiple of 4)
  int g_rw[4], g_ru[4];// ranks of the W and U block of the group
  int g_K[4];          // round4(rw+ru)
  int P_pad;           // total p length (multiple of 4)
  int S1;              // LT row stride
  int S2;              // RT row stride
  int LU;              // lanes per unit (1,2,4); G = 4/LU gates per lane
  int off_LT, off_RT, off_scale, off_bias, off_p, off_vh, off_qmeta;  // float offsets into smem
};

struct WfPlan {
  int L;
  int D;
  int xstride;     // floats per x ring slot
  int off_x;       // x ring
  int off_y;       // y ring
  int off_dense;   // dense kernel (H_last) + bias
  int total_floats;
  WfLayer layers[kMaxLayers];
};

// Layout shared by host (support check + smem size) and device (prologue).
__host__ __device__ inline bool wf_make_plan(const ModelDesc& md, WfPlan& pl) {
  if (md.n_layers > 7) return false;
  if (md.input_dim > 32 || md.n_out > 1) return false;
  pl.L = md.n_layers;
  pl.D = md.input_dim;
  int off = 0;
  for (int l = 0; l < md.n_layers; ++l) {
    const LayerDesc& Ld = md.layers[l];
    WfLayer& w = pl.layers[l];
    w.H = Ld.units;
    w.Din = Ld.d_in;
    if (w.H > 32 || w.Din > 32) return false;
    if (Ld.n_blocks != 2 && Ld.n_blocks != 8) return false;
    w.n_groups = Ld.n_blocks == 2 ? 1 : 4;
    int p = 0, maxK = 4;
    for (int g = 0; g < w.n_groups; ++g) {
      const Block& bw = Ld.blocks[g];
      const Block& bu = Ld.blocks[w.n_groups + g];
      if (bw.rank > 32 || bu.rank > 32 || bw.from_h || !bu.from_h) return false;
      w.g_start[g] = p;
      w.g_rw[g] = bw.rank;
      w.g_ru[g] = bu.rank;
      w.g_K[g] = round4(bw.rank + bu.rank);
      p += w.g_K[g];
      if (w.g_K[g] > maxK) maxK = w.g_K[g];
    }
    for (int g = w.n_groups; g < 4; ++g) w.g_start[g] = w.g_rw[g] = w.g_ru[g] = w.g_K[g] = 0;
    w.P_pad = p;
    if (w.P_pad > 128) return false;
    w.S1 = odd4(w.H > w.Din ? w.H : w.Din);
    w.S2 = odd4(maxK);
    w.LU = w.H > 16 ? 1 : (w.H > 8 ? 2 : 4);
    w.off_LT = off;      off += w.P_pad * w.S1;
    w.off_RT = off;      off += 4 * w.H * w.S2;
    w.off_scale = off;   off += w.P_pad;
    w.off_bias = off;    off += round4(4 * w.H);
    w.off_p = off;       off += w.P_pad;
    w.off_vh = off;      off += 2 * w.S1;   // double-buffered h_l (zero padded to S1)
    w.off_qmeta = off;   off += w.P_pad;    // per-q: bit0 from_h, bits 8.. = padded input length
  }
  pl.xstride = pl.layers[0].S1;
  pl.off_x = off;      off += kXRing * pl.xstride;
  pl.off_y = off;      off += kYRing;
  pl.off_dense = off;  off += 36;
  pl.total_floats = off;
  return (size_t)off * sizeof(float) <= 200 * 1024;
}

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// tanh(x) = 2*sigmoid(2x) - 1 ; absolute error ~1e-7 (ex2.approx + rcp.approx), saturates correctly
__device__ __forceinline__ float fast_tanh(float x) { return fmaf(2.0f, __fdividef(1.0f, 1.0f + __expf(-2.0f * x)), -1.0f); }

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__global__ void __launch_bounds__(288) lstm_wavefront_kernel(const ModelDesc* __restrict__ mdp, ForwardArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ WfPlan pl;
  const ModelDesc& md = *mdp;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  const int B = a.B, T = a.T;
  const bool ret_seq = a.flags & SVDLSTM_RETURN_SEQUENCES;
  const bool backwards = a.flags & SVDLSTM_GO_BACKWARDS;
  const bool time_major = a.flags & SVDLSTM_TIME_MAJOR;

  if (tid == 0) wf_make_plan(md, pl);
  __syncthreads();
  const int L = pl.L, D = pl.D;

  // ---------------- prologue: stage every factor matrix into shared memory -------------------
  for (int i = tid; i < pl.total_floats; i += nthr) smem[i] = 0.f;
  __syncthreads();
  {
    size_t soff = 0;
    for (int l = 0; l < L; ++l) {
      const LayerDesc& Ld = md.layers[l];
      const WfLayer& w = pl.layers[l];
      const int H = w.H;
      // LT + scale + qmeta
      for (int idx = tid; idx < w.P_pad * w.S1; idx += nthr) {
        const int q = idx / w.S1, i = idx - q * w.S1;
        int g = w.n_groups - 1;
        while (g > 0 && q < w.g_start[g]) --g;
        const int kk = q - w.g_start[g];
        float v = 0.f;
        const Block* blk = nullptr;
        int k = 0;
        if (kk < w.g_rw[g]) { blk = &Ld.blocks[g]; k = kk; }
        else if (kk < w.g_rw[g] + w.g_ru[g]) { blk = &Ld.blocks[w.n_groups + g]; k = kk - w.g_rw[g]; }
        if (blk) {
          const int kin = blk->from_h ? H : w.Din;
          if (i < kin) v = blk->left ? blk->left[(size_t)i * blk->left_ld + k] : (i == k ? 1.f : 0.f);
          if (i == 0) {
            smem[w.off_scale + q] = blk->scale ? blk->scale[k] : 1.f;
            smem[w.off_qmeta + q] = __int_as_float((blk->from_h ? 1 : 0) | (round4(kin) << 8));
          }
        }
        smem[w.off_LT + idx] = v;
      }
      // RT + bias
      for (int idx = tid; idx < 4 * H * w.S2; idx += nthr) {
        const int n = idx / w.S2, kk = idx - n * w.S2;
        const int g = w.n_groups == 1 ? 0 : n / H;
        float v = 0.f;
        const Block* blk = nullptr;
        int k = 0;
        if (kk < w.g_rw[g]) { blk = &Ld.blocks[g]; k = kk; }
        else if (kk < w.g_rw[g] + w.g_ru[g]) { blk = &Ld.blocks[w.n_groups + g]; k = kk - w.g_rw[g]; }
        if (blk) {
          int rel = n - blk->out0;
          if (blk->ident) {
            if (rel < blk->rank) v = (rel == k) ? 1.f : 0.f;
            else if (rel - blk->rank < blk->ncols) v = blk->right[(size_t)k * blk->right_ld + (rel - blk->rank)];
          } else if (rel >= 0 && rel < blk->ncols) {
            v = blk->right[(size_t)k * blk->right_ld + rel];
          }
        }
        smem[w.off_RT + idx] = v;
      }
      for (int n = tid; n < 4 * H; n += nthr) smem[w.off_bias + n] = Ld.bias[n];
      // initial h into slot 1 (= slot of timestep -1)
      if (a.h0 != nullptr)
        for (int j = tid; j < H; j += nthr) smem[w.off_vh + w.S1 + j] = a.h0[soff + (size_t)b * H + j];
      soff += (size_t)B * H;
    }
    if (md.n_out > 0) {
      const int HL = pl.layers[L - 1].H;
      for (int j = tid; j < HL; j += nthr) smem[pl.off_dense + j] = md.dense_kernel[j];
      if (tid == 0) smem[pl.off_dense + 32] = md.dense_bias[0];
    }
  }
  __syncthreads();

  // ---------------- per-warp persistent state ------------------------------------------------
  const int role = warp;  // 0: loader, 1..L: layers, L+1: output
  float c_state = 0.f, h_last = 0.f;
  int lyr = role - 1;
  if (role >= 1 && role <= L) {
    if (a.c0 != nullptr) {
      size_t soff = 0;
      for (int l = 0; l < lyr; ++l) soff += (size_t)B * pl.layers[l].H;
      const WfLayer& w0 = pl.layers[lyr];
      const int HP = 32 / w0.LU;
      if (lane < HP && lane < w0.H) {
        c_state = a.c0[soff + (size_t)b * w0.H + lane];
        h_last = a.h0[soff + (size_t)b * w0.H + lane];
      }
    }
  }
  const float* xrow_base = a.x;
  // loader prologue: prefetch the first kPrefetch steps
  if (role == 0) {
    for (int s = 0; s < kPrefetch; ++s) {
      if (s < T && lane < D) {
        const int t = backwards ? (T - 1 - s) : s;
        const float* src = time_major ? xrow_base + ((size_t)t * B + b) * D + lane : xrow_base + ((size_t)b * T + t) * D + lane;
        cp_async4(&smem[pl.off_x + (s % kXRing) * pl.xstride + lane], src);
      }
      cp_async_commit();
    }
    cp_async_wait<0>();
  }
  __syncthreads();

  const int n_ticks = T + L;
  const int HL = pl.layers[L - 1].H;
  const int n_out = md.n_out;
  const int n_y = n_out > 0 ? 1 : HL;
  // register copies of the plan entries this warp touches every tick
  const WfLayer w = pl.layers[(lyr >= 0 && lyr < L) ? lyr : 0];
  const int vin_off = (lyr <= 0) ? pl.off_x : pl.layers[(lyr < L ? lyr : L) - 1].off_vh;
  const int vin_stride = (lyr <= 0) ? pl.xstride : pl.layers[(lyr < L ? lyr : L) - 1].S1;
  const int out_vh_off = pl.layers[L - 1].off_vh, out_vh_stride = pl.layers[L - 1].S1;
  const int off_y = pl.off_y, off_dense = pl.off_dense, off_x = pl.off_x, xstride = pl.xstride;

  for (int tick = 0; tick < n_ticks; ++tick) {
    if (role == 0) {
      // ---- loader: x for step tick+kPrefetch -------------------------------------------------
      const int s = tick + kPrefetch;
      if (s < T && lane < D) {
        const int t = backwards ? (T - 1 - s) : s;
        const float* src = time_major ? xrow_base + ((size_t)t * B + b) * D + lane : xrow_base + ((size_t)b * T + t) * D + lane;
        cp_async4(&smem[off_x + (s % kXRing) * xstride + lane], src);
      }
      cp_async_commit();
      cp_async_wait<kPrefetch - 2>();   // everything up to step tick+1 has landed before the barrier
    } else if (role <= L) {
      const int step = tick - lyr;
      if (step >= 0 && step < T) {
        const int H = w.H;
        const float* vin = &smem[vin_off + ((lyr == 0) ? (step % kXRing) : (step & 1)) * vin_stride];
        const float* vh = &smem[w.off_vh + ((step + 1) & 1) * w.S1];   // h_l(step-1)
        float* vh_out = &smem[w.off_vh + (step & 1) * w.S1];
        float* pbuf = &smem[w.off_p];
        // ---- stage 1 ------------------------------------------------------------------------
        for (int q = lane; q < w.P_pad; q += 32) {
          const int meta = __float_as_int(smem[w.off_qmeta + q]);
          const int kin4 = meta >> 8;
          const float4* v4 = reinterpret_cast<const float4*>((meta & 1) ? vh : vin);
          const float4* l4 = reinterpret_cast<const float4*>(&smem[w.off_LT + q * w.S1]);
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          for (int i = 0; i < kin4 / 4; ++i) {
            const float4 wv = l4[i];
            const float4 vv = v4[i];
            a0 = fmaf(vv.x, wv.x, a0);
            a1 = fmaf(vv.y, wv.y, a1);
            a2 = fmaf(vv.z, wv.z, a2);
            a3 = fmaf(vv.w, wv.w, a3);
          }
          pbuf[q] = ((a0 + a1) + (a2 + a3)) * smem[w.off_scale + q];
        }
        __syncwarp();
        // ---- stage 2 + 3 ----------------------------------------------------------------------
        const int LU = w.LU, G = 4 / LU, HP = 32 / LU;
        const int j = lane % HP, sub = lane / HP;
        float act[4] = {0.f, 0.f, 0.f, 0.f};
        if (j < H) {
          float z[4];
#pragma unroll
          for (int gi = 0; gi < 4; ++gi) {
            if (gi < G) {
              const int gate = sub * G + gi;
              const int n = gate * H + j;
              const int grp = w.n_groups == 1 ? 0 : gate;
              const float4* p4 = reinterpret_cast<const float4*>(pbuf + w.g_start[grp]);
              const float4* r4 = reinterpret_cast<const float4*>(&smem[w.off_RT + n * w.S2]);
              float a0 = smem[w.off_bias + n], a1 = 0.f;
              const int K4 = w.g_K[grp] / 4;
              for (int kk = 0; kk < K4; ++kk) {
                const float4 rv = r4[kk];
                const float4 pv = p4[kk];
                a0 = fmaf(pv.x, rv.x, a0);
                a1 = fmaf(pv.y, rv.y, a1);
                a0 = fmaf(pv.z, rv.z, a0);
                a1 = fmaf(pv.w, rv.w, a1);
              }
              z[gi] = a0 + a1;
            }
          }
#pragma unroll
          for (int gi = 0; gi < 4; ++gi) {
            if (gi < G) {
              const int gate = sub * G + gi;
              act[gi] = (gate == 2) ? fast_tanh(z[gi]) : fast_sigmoid(z[gi]);
            }
          }
        }
        float ig, fg, gg, og;
        if (LU == 1) {
          ig = act[0]; fg = act[1]; gg = act[2]; og = act[3];
        } else if (LU == 2) {
          const float o0 = __shfl_xor_sync(0xffffffffu, act[0], 16);
          const float o1 = __shfl_xor_sync(0xffffffffu, act[1], 16);
          ig = act[0]; fg = act[1]; gg = o0; og = o1;   // valid in sub==0 lanes
        } else {
          const float v1 = __shfl_sync(0xffffffffu, act[0], (lane & 7) + 8);
          const float v2 = __shfl_sync(0xffffffffu, act[0], (lane & 7) + 16);
          const float v3 = __shfl_sync(0xffffffffu, act[0], (lane & 7) + 24);
          ig = act[0]; fg = v1; gg = v2; og = v3;
        }
        if (sub == 0 && j < H) {
          c_state = fmaf(fg, c_state, ig * gg);
          h_last = og * fast_tanh(c_state);
          vh_out[j] = h_last;
        }
      }
    } else if (role == L + 1) {
      // ---- output stage ----------------------------------------------------------------------
      const int step = tick - L;
      if (step >= 0 && step < T) {
        const float* hv = &smem[out_vh_off + (step & 1) * out_vh_stride];
        if (n_out > 0) {
          float v = (lane < HL) ? hv[lane] * smem[off_dense + lane] : 0.f;
#pragma unroll
          for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
          v += smem[off_dense + 32];
          if (ret_seq) {
            if (lane == 0) smem[off_y + (step % kYRing)] = v;
            if ((step % kYRing) == kYRing - 1 || step == T - 1) {
              __syncwarp();
              const int first = step - (step % kYRing);
              if (first + lane <= step) {
                const size_t yi = time_major ? ((size_t)(first + lane) * B + b) : ((size_t)b * T + first + lane);
                a.y[yi] = smem[off_y + lane];
              }
              __syncwarp();
            }
          } else if (step == T - 1 && lane == 0) {
            a.y[b] = v;
          }
        } else if (ret_seq || step == T - 1) {
          if (lane < HL) {
            size_t yi;
            if (!ret_seq) yi = (size_t)b * n_y + lane;
            else yi = time_major ? ((size_t)step * B + b) * n_y + lane : ((size_t)b * T + step) * n_y + lane;
            a.y[yi] = hv[lane];
          }
        }
      }
    }
    __syncthreads();
  }

  // ---------------- final state ----------------------------------------------------------------
  if (role >= 1 && role <= L && (a.h_n || a.c_n)) {
    size_t soff = 0;
    for (int l = 0; l < lyr; ++l) soff += (size_t)B * pl.layers[l].H;
    const int HP = 32 / w.LU;
    if (lane < HP && lane < w.H) {
      if (a.h_n) a.h_n[soff + (size_t)b * w.H + lane] = h_last;
      if (a.c_n) a.c_n[soff + (size_t)b * w.H + lane] = c_state;
    }
  }
}

}  // namespace

bool wavefront_supported(const ModelDesc& md, const ForwardArgs& a) {
  if (a.mask != nullptr) return false;
  WfPlan pl;
  return wf_make_plan(md, pl);
}

int run_wavefront(const ModelDesc& md, const ModelDesc* dev_md, const ForwardArgs& a, cudaStream_t stream, int* launches) {
  WfPlan pl;
  SVD_REQUIRE(wf_make_plan(md, pl), "wavefront engine: model does not fit (units/input_dim/ranks <= 32, <= 7 layers, n_out <= 1)");
  const size_t smem = (size_t)pl.total_floats * sizeof(float);
  SVD_CUDA_TRY(cudaFuncSetAttribute(lstm_wavefront_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = 32 * (md.n_layers + 2);
  lstm_wavefront_kernel<<<a.B, threads, smem, stream>>>(dev_md, a);
  SVD_CUDA_TRY(cudaGetLastError());
  *launches = 1;
  return 0;
}

}  // namespace svdlstm
