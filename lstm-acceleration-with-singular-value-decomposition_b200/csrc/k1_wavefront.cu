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
#include <stdlib.h>

#include "common.cuh"

namespace svdlstm {

namespace {

// One CTA barrier costs about as much as one layer-step's whole dependent chain (measured: ~180 of a 498-cycle tick), so a
// wavefront tick covers kSteps consecutive timesteps per layer: the recurrence inside a tick needs only a __syncwarp (a layer
// is one warp), and the barrier is paid once per kSteps.
constexpr int kSteps = 4;       // timesteps per wavefront tick and layer
constexpr int kHRing = 2 * kSteps;   // h_l ring: the half being written this tick + the half the next layer reads
constexpr int kXRing = 32;      // x_t ring depth (steps)
constexpr int kPrefetch = 5;    // cp.async distance (ticks of kSteps steps): (kPrefetch + 1) * kSteps <= kXRing
constexpr int kYRing = 32;      // y staging before a coalesced store
static_assert((kPrefetch + 1) * kSteps <= kXRing && kYRing % kSteps == 0, "ring sizes");

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
  int WS, off_we;      // dense mode: row stride and offset of W_eff [4H][WS] = [input part | recurrent part], each WS/2 wide
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
__host__ __device__ inline bool wf_make_plan(const ModelDesc& md, WfPlan& pl, bool dense = false) {
  if (md.n_layers > 6) return false;
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
    w.off_p = off;       off += w.P_pad + 64;   // + zero tail: stage 2 reads a full K4*4 window
    w.off_vh = off;      off += kHRing * w.S1;   // ring of h_l(t) slots, slot = t % kHRing (zero padded to S1)
    w.off_qmeta = off;   off += w.P_pad;    // per-q: bit0 from_h, bits 8.. = padded input length
    w.WS = 2 * round4(w.H > w.Din ? w.H : w.Din);
    w.off_we = off;      off += dense ? 4 * w.H * w.WS : 0;
  }
  pl.xstride = pl.layers[0].S1;
  pl.off_x = off;      off += kXRing * pl.xstride;
  pl.off_y = off;      off += kYRing;
  pl.off_dense = off;  off += 36;
  pl.total_floats = off;
  return (size_t)off * sizeof(float) <= 200 * 1024;
}

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid(x) = 1/(1+2^(-x log2 e)); tanh(x) = 2*sigmoid(2x) - 1.  Two MUFU ops each, absolute error
// ~1e-7 (ex2.approx / rcp.approx are <= 2 ulp); both saturate correctly (2^-inf = 0, rcp(inf) = 0).
__device__ __forceinline__ float fast_sigmoid(float x) { return rcp_ftz(1.0f + ex2_ftz(-1.4426950408889634f * x)); }
__device__ __forceinline__ float fast_tanh(float x) { return fmaf(2.0f, rcp_ftz(1.0f + ex2_ftz(-2.8853900817779268f * x)), -1.0f); }

// explicit shared-space accesses on precomputed 32-bit addresses (no generic->shared conversion per tick)
__device__ __forceinline__ float4 lds128(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(unsigned addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
// Blackwell packed FP32x2 FMA: two IEEE fma.rn per issue slot (SASS FFMA2)
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pack2(float lo, float hi) {
  f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2_t ffma2(f2_t a, f2_t b, f2_t c) {
  f2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Template shape (max over layers, zero padded): KIN4 float4s of stage-1 input, MP stage-1 columns
// per lane, K4 float4s of stage-2 contraction, G gates per lane (2: two lanes per unit, 4: one).
//
// STREAM = true is the real-time service (svd_classes_v3.py:421-426 `stateful`, driven at the reference's one-sample-every-400-us
// setting, train_full_model_v4.py:14-16): ONE persistent CTA that never returns to the host between samples.  Its loader warp
// polls a ring in host-mapped pinned memory for the next sample, the layer warps run that step one after the other (sub-tick l =
// layer l: the prediction of sample s must not wait for sample s+L, so the wavefront is NOT used across samples), and the
// output warp writes the prediction straight into host-mapped memory.  No launch, no memcpy, no stream synchronisation per
// sample; state (h, c) stays in registers / shared memory and is parked in device memory when the kernel leaves (host `stop`,
// or no sample for `idle_ns`: a forgotten stream must never pin an SM -- or block a cudaDeviceSynchronize -- for ever).
//
// DENSE = true (small models: units, input_dim <= 16): the latency regime has nothing to gain from the factored evaluation order
// -- at H = 15 one fused contraction [x_t | h(t-1)] . W_eff with W_eff = (L sigma) R costs 900 MACs against 1 125 for the two thin
// ones at full rank, and, more to the point, removes one FFMA chain and one shared-memory round trip (p) from the dependent
// chain of every tick.  W_eff is formed ONCE per launch in the prologue (float64 accumulation, so the 2-factor [I | C] forms
// lose nothing to cancellation) and lives in registers; rank truncation is exactly preserved (W_eff has rank r).
template <int KIN4, int MP, int K4, int G, bool STREAM = false, bool DENSE = false>
__global__ void __launch_bounds__(256, 1) lstm_wavefront_kernel(const ModelDesc* __restrict__ mdp, ForwardArgs a, StreamArgs sa) {
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

  if (tid == 0) wf_make_plan(md, pl, DENSE);
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
        for (int j = tid; j < H; j += nthr) {
          for (int sl = 0; sl < kHRing; ++sl)      // h(-1) lives in slot kHRing-1; a resumed stream may start on any step
            if (STREAM || sl == kHRing - 1) smem[w.off_vh + sl * w.S1 + j] = a.h0[soff + (size_t)b * H + j];
        }
      soff += (size_t)B * H;
    }
    if (md.n_out > 0) {
      const int HL = pl.layers[L - 1].H;
      for (int j = tid; j < HL; j += nthr) smem[pl.off_dense + j] = md.dense_kernel[j];
      if (tid == 0) smem[pl.off_dense + 32] = md.dense_bias[0];
    }
  }
  __syncthreads();
  if constexpr (DENSE) {
    for (int l = 0; l < L; ++l) {
      const WfLayer& w = pl.layers[l];
      const int half = w.WS / 2;
      for (int idx = tid; idx < 4 * w.H * w.WS; idx += nthr) {
        const int n = idx / w.WS, c = idx - n * w.WS;
        const int fh = c >= half ? 1 : 0, i = c - fh * half;
        const int grp = w.n_groups == 1 ? 0 : n / w.H;
        double acc = 0.0;
        for (int kk = 0; kk < w.g_K[grp]; ++kk) {
          const int q = w.g_start[grp] + kk;
          if ((__float_as_int(smem[w.off_qmeta + q]) & 1) != fh) continue;
          acc += (double)smem[w.off_RT + n * w.S2 + kk] * (double)smem[w.off_scale + q] * (double)smem[w.off_LT + q * w.S1 + i];
        }
        smem[w.off_we + idx] = (float)acc;
      }
    }
    __syncthreads();
  }

  // ---------------- per-warp persistent state: weights -> REGISTERS ---------------------------
  const int role = warp;  // 0: loader, 1..L: layers, L+1: output
  const int lyr = role - 1;
  const bool is_layer = role >= 1 && role <= L;
  const WfLayer& wl = pl.layers[is_layer ? lyr : 0];
  const int H = wl.H;
  constexpr int HP = (G == 4) ? 32 : 16;
  const int j = lane % HP, sub = lane / HP;

  // stage-1 weights of my columns over the concatenated input [x_t | h(t-1)] (packed pairs; the half a
  // column does not use is zero, so no per-lane select is needed; sigma is folded in: v.(L*sigma))
  f2_t w1[DENSE ? 1 : MP][DENSE ? 1 : KIN4 * 4];
  f2_t w2[G][DENSE ? 1 : K4 * 2];       // stage-2 weights of my gate columns (packed pairs)
  f2_t wd[G][DENSE ? KIN4 * 4 : 1];     // dense mode: W_eff columns of my gates over [x_t | h(t-1)] (packed pairs)
  float bias2[G];
  unsigned pofs[G];         // shared address of the p window each of my gates contracts with
  const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
  float c_state = 0.f, h_last = 0.f;
  if (is_layer && DENSE) {
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
      const int gate = sub * G + gi;
      const bool ok = j < H;
      const int n = gate * H + (ok ? j : 0);
      const int half = wl.WS / 2;
      pofs[gi] = 0;
      bias2[gi] = ok ? smem[wl.off_bias + n] : 0.f;
#pragma unroll
      for (int i = 0; i < KIN4 * 2; ++i) {
        const float* row = smem + wl.off_we + n * wl.WS;
        wd[gi][i] = pack2((ok && 2 * i < half) ? row[2 * i] : 0.f, (ok && 2 * i + 1 < half) ? row[2 * i + 1] : 0.f);
        wd[gi][KIN4 * 2 + i] = pack2((ok && 2 * i < half) ? row[half + 2 * i] : 0.f, (ok && 2 * i + 1 < half) ? row[half + 2 * i + 1] : 0.f);
      }
    }
  }
  if (is_layer && !DENSE) {
#pragma unroll
    for (int m = 0; m < MP; ++m) {
      const int q = lane + 32 * m;
      const bool ok = q < wl.P_pad;
      const int meta = ok ? __float_as_int(smem[wl.off_qmeta + q]) : 0;
      const bool fh = meta & 1;
      const float sc = ok ? smem[wl.off_scale + q] : 0.f;
#pragma unroll
      for (int i = 0; i < KIN4 * 2; ++i) {
        const float lo = (ok && 2 * i < wl.S1) ? sc * smem[wl.off_LT + q * wl.S1 + 2 * i] : 0.f;
        const float hi = (ok && 2 * i + 1 < wl.S1) ? sc * smem[wl.off_LT + q * wl.S1 + 2 * i + 1] : 0.f;
        w1[m][i] = fh ? 0ull : pack2(lo, hi);
        w1[m][KIN4 * 2 + i] = fh ? pack2(lo, hi) : 0ull;
      }
    }
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
      const int gate = sub * G + gi;
      const bool ok = j < H;
      const int n = gate * H + (ok ? j : 0);
      const int grp = wl.n_groups == 1 ? 0 : gate;
      pofs[gi] = smem_base + 4u * (unsigned)(wl.off_p + wl.g_start[grp]);
      bias2[gi] = ok ? smem[wl.off_bias + n] : 0.f;
#pragma unroll
      for (int kk = 0; kk < K4 * 2; ++kk) {
        const float lo = (ok && 2 * kk < wl.g_K[grp]) ? smem[wl.off_RT + n * wl.S2 + 2 * kk] : 0.f;
        const float hi = (ok && 2 * kk + 1 < wl.g_K[grp]) ? smem[wl.off_RT + n * wl.S2 + 2 * kk + 1] : 0.f;
        w2[gi][kk] = pack2(lo, hi);
      }
    }
  }
  if (is_layer) {
    if (a.c0 != nullptr) {
      size_t soff = 0;
      for (int l = 0; l < lyr; ++l) soff += (size_t)B * pl.layers[l].H;
      if (sub == 0 && j < H) {
        c_state = a.c0[soff + (size_t)b * H + j];
        h_last = a.h0[soff + (size_t)b * H + j];
      }
    }
  }
  const float* xrow_base = a.x;
  const int off_x = pl.off_x, xstride = pl.xstride;
  // loader prologue: prefetch the first kPrefetch ticks (kSteps steps each, one cp.async group per tick)
  if (!STREAM && role == 0) {
    for (int tk = 0; tk < kPrefetch; ++tk) {
      for (int u = 0; u < kSteps; ++u) {
        const int s = tk * kSteps + u;
        if (s < T && lane < D) {
          const int t = backwards ? (T - 1 - s) : s;
          const float* src = time_major ? xrow_base + ((size_t)t * B + b) * D + lane : xrow_base + ((size_t)b * T + t) * D + lane;
          cp_async4(&smem[off_x + (s % kXRing) * xstride + lane], src);
        }
      }
      cp_async_commit();
    }
    cp_async_wait<0>();
  }
  __syncthreads();

  const int n_ticks = (T + kSteps - 1) / kSteps + L;
  const int HL = pl.layers[L - 1].H;
  const int n_out = md.n_out;
  const int n_y = n_out > 0 ? 1 : HL;
  const int vin_off = (lyr <= 0 || !is_layer) ? off_x : pl.layers[lyr - 1].off_vh;
  const int vin_stride = (lyr <= 0 || !is_layer) ? xstride : pl.layers[lyr - 1].S1;
  const int my_Ppad = wl.P_pad;
  const unsigned vin_addr = smem_base + 4u * (unsigned)vin_off, vin_sb = 4u * (unsigned)vin_stride;
  const unsigned vh_addr = smem_base + 4u * (unsigned)wl.off_vh, vh_sb = 4u * (unsigned)wl.S1;
  const unsigned p_st_addr = smem_base + 4u * (unsigned)(wl.off_p + lane);
  const bool same_window = __all_sync(0xffffffffu, pofs[0] == pofs[G - 1]);   // warp-uniform
  const int out_vh_off = pl.layers[L - 1].off_vh, out_vh_stride = pl.layers[L - 1].S1;
  const int off_y = pl.off_y, off_dense = pl.off_dense;
  const float dense_w = (role == L + 1 && n_out > 0 && lane < HL) ? smem[off_dense + lane] : 0.f;
  const float dense_b = (n_out > 0) ? smem[off_dense + 32] : 0.f;

  // dense mode, input side of a step: zx = bias + x_t . W_eff[input part].  No recurrence: the wavefront tick computes it for all
  // of its kSteps steps up front (pure issue-bound FFMA2 work), which takes it off the dependent chain of every step.
  auto dense_x = [&](const int step, const int xslot, float (&zx)[G]) {
        const unsigned va = vin_addr + (unsigned)((lyr == 0) ? xslot : (step % kHRing)) * vin_sb;
        float4 vi[KIN4];
#pragma unroll
        for (int i = 0; i < KIN4; ++i) vi[i] = lds128(va + 16u * i);
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
          f2_t a0 = pack2(bias2[gi], 0.f), a1 = 0ull;
#pragma unroll
          for (int i = 0; i < KIN4; ++i) {
            a0 = ffma2(pack2(vi[i].x, vi[i].y), wd[gi][2 * i + 0], a0);
            a1 = ffma2(pack2(vi[i].z, vi[i].w), wd[gi][2 * i + 1], a1);
          }
          float s0, s1, s2, s3;
          unpack2(a0, s0, s1);
          unpack2(a1, s2, s3);
          zx[gi] = (s0 + s1) + (s2 + s3);
        }
  };
  // one step of the layer this warp owns: `step` selects the h ring slots, `xslot` the x ring slot (layer 0); dense mode: zx = the
  // input side of this step from dense_x
  auto layer_step = [&](const int step, const int xslot, const float (&zx)[G]) {
        const unsigned va = vin_addr + (unsigned)((lyr == 0) ? xslot : (step % kHRing)) * vin_sb;
        const unsigned ha = vh_addr + (unsigned)((step + kHRing - 1) % kHRing) * vh_sb;       // h_l(step-1)
        const unsigned ho = vh_addr + (unsigned)(step % kHRing) * vh_sb + 4u * (unsigned)j;
        float zz[G];
        if constexpr (DENSE) {
          // ---- recurrent side: z = zx + h(t-1) . W_eff[recurrent part], two independent FFMA2 chains of depth KIN4 per gate ----
          float4 vh[KIN4];
#pragma unroll
          for (int i = 0; i < KIN4; ++i) vh[i] = lds128(ha + 16u * i);
#pragma unroll
          for (int gi = 0; gi < G; ++gi) {
            f2_t b0 = pack2(zx[gi], 0.f), b1 = 0ull;
#pragma unroll
            for (int i = 0; i < KIN4; ++i) {
              b0 = ffma2(pack2(vh[i].x, vh[i].y), wd[gi][KIN4 * 2 + 2 * i + 0], b0);
              b1 = ffma2(pack2(vh[i].z, vh[i].w), wd[gi][KIN4 * 2 + 2 * i + 1], b1);
            }
            float u0, u1, u2, u3;
            unpack2(b0, u0, u1);
            unpack2(b1, u2, u3);
            zz[gi] = (u0 + u1) + (u2 + u3);
          }
        } else {
        // ---- stage 1: p[q] = scale * <v, LT[q,:]> ----------------------------------------------
        float4 vi[KIN4], vh[KIN4];
#pragma unroll
        for (int i = 0; i < KIN4; ++i) {
          vi[i] = lds128(va + 16u * i);
          vh[i] = lds128(ha + 16u * i);
        }
#pragma unroll
        for (int m = 0; m < MP; ++m) {
          f2_t a0 = 0ull, a1 = 0ull;
#pragma unroll
          for (int i = 0; i < KIN4; ++i) {
            a0 = ffma2(pack2(vi[i].x, vi[i].y), w1[m][2 * i + 0], a0);
            a1 = ffma2(pack2(vi[i].z, vi[i].w), w1[m][2 * i + 1], a1);
          }
#pragma unroll
          for (int i = 0; i < KIN4; ++i) {
            a0 = ffma2(pack2(vh[i].x, vh[i].y), w1[m][KIN4 * 2 + 2 * i + 0], a0);
            a1 = ffma2(pack2(vh[i].z, vh[i].w), w1[m][KIN4 * 2 + 2 * i + 1], a1);
          }
          float s0, s1, s2, s3;
          unpack2(a0, s0, s1);
          unpack2(a1, s2, s3);
          if (lane + 32 * m < my_Ppad) sts32(p_st_addr + 128u * m, (s0 + s1) + (s2 + s3));
        }
        __syncwarp();
        // ---- stage 2: z = bias + <p[window], RT[n,:]> ; stage 3: gates ---------------------------
        if (same_window) {   // merged / full forms: every gate contracts with the same p window
          float4 pw[K4];
#pragma unroll
          for (int kk = 0; kk < K4; ++kk) pw[kk] = lds128(pofs[0] + 16u * kk);
#pragma unroll
          for (int gi = 0; gi < G; ++gi) {
            f2_t a0 = pack2(bias2[gi], 0.f), a1 = 0ull;
#pragma unroll
            for (int kk = 0; kk < K4; ++kk) {
              a0 = ffma2(pack2(pw[kk].x, pw[kk].y), w2[gi][2 * kk + 0], a0);
              a1 = ffma2(pack2(pw[kk].z, pw[kk].w), w2[gi][2 * kk + 1], a1);
            }
            float s0, s1, s2, s3;
            unpack2(a0, s0, s1);
            unpack2(a1, s2, s3);
            zz[gi] = (s0 + s1) + (s2 + s3);
          }
        } else {             // split forms: one p window per gate
#pragma unroll
          for (int gi = 0; gi < G; ++gi) {
            float4 pw[K4];
#pragma unroll
            for (int kk = 0; kk < K4; ++kk) pw[kk] = lds128(pofs[gi] + 16u * kk);
            f2_t a0 = pack2(bias2[gi], 0.f), a1 = 0ull;
#pragma unroll
            for (int kk = 0; kk < K4; ++kk) {
              a0 = ffma2(pack2(pw[kk].x, pw[kk].y), w2[gi][2 * kk + 0], a0);
              a1 = ffma2(pack2(pw[kk].z, pw[kk].w), w2[gi][2 * kk + 1], a1);
            }
            float s0, s1, s2, s3;
            unpack2(a0, s0, s1);
            unpack2(a1, s2, s3);
            zz[gi] = (s0 + s1) + (s2 + s3);
          }
        }
        }
        float act[G];
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
          // sigmoid for i,f,o ; tanh for the candidate gate (gate 2): same code, scale/offset select
          const bool is_tanh = (G == 4) ? (gi == 2) : (sub == 1 && gi == 0);
          const float k = is_tanh ? -2.8853900817779268f : -1.4426950408889634f;
          const float r = rcp_ftz(1.0f + ex2_ftz(k * zz[gi]));
          act[gi] = is_tanh ? fmaf(2.0f, r, -1.0f) : r;
        }
        float ig, fg, gg, og;
        if (G == 4) {
          ig = act[0]; fg = act[1]; gg = act[2]; og = act[3];
        } else {
          const float o0 = __shfl_xor_sync(0xffffffffu, act[0], 16);
          const float o1 = __shfl_xor_sync(0xffffffffu, act[1], 16);
          ig = act[0]; fg = act[1]; gg = o0; og = o1;   // valid in sub==0 lanes
        }
        c_state = fmaf(fg, c_state, ig * gg);
        h_last = og * fast_tanh(c_state);
        if (sub == 0 && j < H) sts32(ho, h_last);
  };
  auto output_step = [&](const int step) {
        const float* hv = &smem[out_vh_off + (step % kHRing) * out_vh_stride];
        if (n_out > 0) {
          float v = (lane < HL) ? hv[lane] * dense_w : 0.f;
#pragma unroll
          for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
          v += dense_b;
          if constexpr (STREAM) {
            // [tag, y] in one 64-byte line of host memory, one store instruction: the host sees the tag only with the value
            if (lane < 2) sa.out[(uint32_t)step % kStreamSlots].w[lane] = lane == 0 ? (uint32_t)step + 1u : __float_as_uint(v);
          } else if (ret_seq) {
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
        } else if constexpr (STREAM) {
          if (lane < 16) sa.out[(uint32_t)step % kStreamSlots].w[lane] = lane == 0 ? (uint32_t)step + 1u : (lane - 1 < HL ? __float_as_uint(hv[lane - 1]) : 0u);
        } else if (ret_seq || step == T - 1) {
          if (lane < HL) {
            size_t yi;
            if (!ret_seq) yi = (size_t)b * n_y + lane;
            else yi = time_major ? ((size_t)step * B + b) * n_y + lane : ((size_t)b * T + step) * n_y + lane;
            a.y[yi] = hv[lane];
          }
        }
  };

  if constexpr (!STREAM) {
    for (int tick = 0; tick < n_ticks; ++tick) {
      if (role == 0) {
        // ---- loader: x for the steps of tick + kPrefetch -----------------------------------------
#pragma unroll
        for (int u = 0; u < kSteps; ++u) {
          const int s = (tick + kPrefetch) * kSteps + u;
          if (s < T && lane < D) {
            const int t = backwards ? (T - 1 - s) : s;
            const float* src = time_major ? xrow_base + ((size_t)t * B + b) * D + lane : xrow_base + ((size_t)b * T + t) * D + lane;
            cp_async4(&smem[off_x + (s % kXRing) * xstride + lane], src);
          }
        }
        cp_async_commit();
        cp_async_wait<kPrefetch - 2>();   // everything up to the steps of tick+1 has landed before the barrier
      } else if (is_layer) {
        const int s0 = (tick - lyr) * kSteps;
        if (s0 >= 0) {
#pragma unroll 1
          float zx[kSteps][G];
          if constexpr (DENSE) {
#pragma unroll
            for (int u = 0; u < kSteps; ++u) dense_x(s0 + u, (s0 + u) % kXRing, zx[u]);   // (steps >= T read stale ring slots: unused)
          }
#pragma unroll
          for (int u = 0; u < kSteps; ++u) {
            const int step = s0 + u;
            if (step < T) {
              layer_step(step, step % kXRing, zx[u]);
              __syncwarp();     // h_l(step) written by this warp's lanes is read by all of them in the next step
            }
          }
        }
      } else if (role == L + 1) {
        const int s0 = (tick - L) * kSteps;
        if (s0 >= 0) {
#pragma unroll 1
          for (int u = 0; u < kSteps; ++u) {
            const int step = s0 + u;
            if (step >= T) break;
            output_step(step);
          }
        }
      }
      __syncthreads();
    }
  } else {
    // ======================= real-time service loop =============================================
    __shared__ int s_exit;
    if (tid == 0) s_exit = 0;
    __syncthreads();
    uint32_t seq = sa.first_seq;     // samples consumed so far (by every launch of this stream)
    while (true) {
      if (role == 0) {
        // ---- poll the host-mapped input ring for sample `seq` (tag seq+1 in BOTH 64-byte lines of its slot) ----
        const uint32_t want = (seq & 0x7fffffffu) + 1u;
        const volatile uint32_t* slot = sa.in[seq % kStreamSlots].w;
        unsigned long long t0 = 0;
        uint32_t v = 0;
        int polls = 0, reason = 0;
        while (true) {
          v = slot[lane];                          // ONE 128-byte read over PCIe per poll
          const uint32_t tag_a = __shfl_sync(0xffffffffu, v, 0), tag_b = __shfl_sync(0xffffffffu, v, 16);
          if (tag_a == want && tag_b == want) break;
          if ((++polls & 7) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            int r = 0;
            if (lane == 0) {
              if (sa.ctl->stop != 0u) r = 1;
              else if (now - t0 > sa.idle_ns) r = 2;
            }
            reason = __shfl_sync(0xffffffffu, r, 0);
            if (reason) break;
          }
        }
        if (reason) {
          if (lane == 0) s_exit = reason;
        } else {
          // line A = [tag, x0..x14], line B = [tag, x15..x29]
          const int xi = lane < 16 ? lane - 1 : lane - 2;
          if (lane != 0 && lane != 16 && xi < D) smem[off_x + xi] = __uint_as_float(v);
        }
      }
      __syncthreads();
      if (s_exit) break;
      const int step = (int)(seq & 0x7fffffffu);
      for (int sub = 0; sub <= L; ++sub) {
        if (is_layer && lyr == sub) {
          float zx[G];
          if constexpr (DENSE) dense_x(step, 0, zx);
          layer_step(step, 0, zx);
        }
        else if (role == L + 1 && sub == L) output_step(step);
        __syncthreads();
      }
      ++seq;
    }
    if (tid == 0) {
      sa.ctl->consumed = seq;
      sa.ctl->exit_reason = (uint32_t)s_exit;
      __threadfence_system();
      sa.ctl->exited = sa.generation;
    }
  }

  // ---------------- final state ----------------------------------------------------------------
  if (is_layer && (a.h_n || a.c_n)) {
    size_t soff = 0;
    for (int l = 0; l < lyr; ++l) soff += (size_t)B * pl.layers[l].H;
    if (sub == 0 && j < H) {
      if (a.h_n) a.h_n[soff + (size_t)b * H + j] = h_last;
      if (a.c_n) a.c_n[soff + (size_t)b * H + j] = c_state;
    }
  }
}

// ---- template selection ------------------------------------------------------------------------
struct WfShape {
  int kin4, mp, k4, g;
};

inline bool wf_shape(const WfPlan& pl, WfShape& s) {
  int kin = 0, pp = 0, k = 0, hmax = 0;
  for (int l = 0; l < pl.L; ++l) {
    const WfLayer& w = pl.layers[l];
    kin = kin > round4(w.H > w.Din ? w.H : w.Din) ? kin : round4(w.H > w.Din ? w.H : w.Din);
    pp = pp > w.P_pad ? pp : w.P_pad;
    for (int g = 0; g < w.n_groups; ++g) k = k > w.g_K[g] ? k : w.g_K[g];
    hmax = hmax > w.H ? hmax : w.H;
  }
  s.kin4 = kin <= 16 ? 4 : 8;
  s.mp = pp <= 32 ? 1 : (pp <= 64 ? 2 : 4);
  s.k4 = k <= 8 ? 2 : (k <= 16 ? 4 : (k <= 32 ? 8 : 16));
  s.g = hmax <= 16 ? 2 : 4;
  // register budget of the weight arrays (floats per lane)
  return s.mp * s.kin4 * 8 + s.g * s.k4 * 4 <= 200;
}

typedef void (*WfKernel)(const ModelDesc*, ForwardArgs, StreamArgs);

template <int KIN4, int MP, int K4, bool STREAM>
WfKernel wf_pick_g(int g) {
  if (g == 2) return lstm_wavefront_kernel<KIN4, MP, K4, 2, STREAM>;
  if constexpr (MP * KIN4 * 8 + 4 * K4 * 4 <= 200) return lstm_wavefront_kernel<KIN4, MP, K4, 4, STREAM>;
  return nullptr;
}
template <int KIN4, int MP, bool STREAM>
WfKernel wf_pick_k(int k4, int g) {
  switch (k4) {
    case 2: return wf_pick_g<KIN4, MP, 2, STREAM>(g);
    case 4: return wf_pick_g<KIN4, MP, 4, STREAM>(g);
    case 8: return wf_pick_g<KIN4, MP, 8, STREAM>(g);
    default:
      if constexpr (MP * KIN4 * 8 + 2 * 16 * 4 <= 200) return wf_pick_g<KIN4, MP, 16, STREAM>(g);
      return nullptr;
  }
}
template <int KIN4, bool STREAM>
WfKernel wf_pick_mp(int mp, int k4, int g) {
  switch (mp) {
    case 1: return wf_pick_k<KIN4, 1, STREAM>(k4, g);
    case 2: return wf_pick_k<KIN4, 2, STREAM>(k4, g);
    default: return wf_pick_k<KIN4, 4, STREAM>(k4, g);
  }
}
template <bool STREAM = false>
inline WfKernel wf_pick(const WfShape& s) { return s.kin4 == 4 ? wf_pick_mp<4, STREAM>(s.mp, s.k4, s.g) : wf_pick_mp<8, STREAM>(s.mp, s.k4, s.g); }

// Dense mode (one fused contraction per layer-tick, W_eff in registers) for the small-model latency regime: units and input
// width <= 16 (G = 2, KIN4 = 4: 64 weight registers per lane).  SVDLSTM_WF_FACTORED=1 keeps the two-stage factored evaluation.
inline bool wf_dense(const WfShape& s) {
  const char* e = getenv("SVDLSTM_WF_FACTORED");     // read per call: bench.py times both evaluation orders in one process
  const bool off = e && e[0] && e[0] != '0';
  return !off && s.kin4 == 4 && s.g == 2;
}
template <bool STREAM>
inline WfKernel wf_pick_dense() { return lstm_wavefront_kernel<4, 1, 2, 2, STREAM, true>; }

}  // namespace

bool wavefront_supported(const ModelDesc& md, const ForwardArgs& a) {
  if (a.mask != nullptr) return false;
  WfPlan pl;
  if (!wf_make_plan(md, pl)) return false;
  WfShape sh;
  return wf_shape(pl, sh) && wf_pick(sh) != nullptr;
}

int run_wavefront(const ModelDesc& md, const ModelDesc* dev_md, const ForwardArgs& a, cudaStream_t stream, int* launches) {
  WfPlan pl;
  SVD_REQUIRE(wf_make_plan(md, pl), "wavefront engine: model does not fit (units/input_dim/ranks <= 32, <= 7 layers, n_out <= 1)");
  WfShape sh;
  WfKernel kern = wf_shape(pl, sh) ? wf_pick(sh) : nullptr;
  SVD_REQUIRE(kern != nullptr, "wavefront engine: factor shapes exceed the register-resident budget");
  if (wf_dense(sh)) {
    WfPlan pd;
    if (wf_make_plan(md, pd, true)) {
      pl = pd;
      kern = wf_pick_dense<false>();
    }
  }
  const size_t smem = (size_t)pl.total_floats * sizeof(float);
  SVD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = 32 * (md.n_layers + 2);
  kern<<<a.B, threads, smem, stream>>>(dev_md, a, StreamArgs{});
  SVD_CUDA_TRY(cudaGetLastError());
  *launches = 1;
  return 0;
}

// The real-time service kernel (one persistent CTA): see stream.cu for the host half of the protocol.
bool stream_supported(const ModelDesc& md, const char** why) {
  ForwardArgs a{};
  if (!wavefront_supported(md, a)) { *why = "the real-time stream runs on the wavefront engine: units, input_dim, ranks <= 32, <= 6 layers, n_out <= 1"; return false; }
  if (md.input_dim > 30) { *why = "the real-time stream carries at most 30 input features per sample (one 128-byte ring slot)"; return false; }
  if (md.n_out == 0 && md.layers[md.n_layers - 1].units > 15) { *why = "the real-time stream returns at most 15 values per sample (one 64-byte ring slot)"; return false; }
  return true;
}

int launch_wavefront_stream(const ModelDesc& md, const ModelDesc* dev_md, float* state_h, float* state_c, const StreamArgs& sa, cudaStream_t stream) {
  WfPlan pl;
  SVD_REQUIRE(wf_make_plan(md, pl), "real-time stream: model does not fit the wavefront engine");
  WfShape sh;
  WfKernel kern = wf_shape(pl, sh) ? wf_pick<true>(sh) : nullptr;
  SVD_REQUIRE(kern != nullptr, "real-time stream: factor shapes exceed the register-resident budget");
  if (wf_dense(sh)) {
    WfPlan pd;
    if (wf_make_plan(md, pd, true)) {
      pl = pd;
      kern = wf_pick_dense<true>();
    }
  }
  const size_t smem = (size_t)pl.total_floats * sizeof(float);
  SVD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ForwardArgs a{nullptr, nullptr, state_h, state_c, state_h, state_c, nullptr, 1, 1, SVDLSTM_RETURN_SEQUENCES};
  kern<<<1, 32 * (md.n_layers + 2), smem, stream>>>(dev_md, a, sa);
  SVD_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace svdlstm
