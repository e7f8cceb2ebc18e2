// K1 (general): FP32 CUDA-core persistent recurrent kernel for ANY cell form / shape.
//
// One CTA owns BT sequences for all T steps and all layers (sequences are independent, so no
// grid-level synchronisation exists anywhere).  Recurrent state (h, c per layer) lives in shared
// memory for the whole launch; per step and layer the CTA runs
//   stage 1  p = scale * (in . left)            thread per intermediate column, BT accumulators
//   stage 2  z = bias + sum_blocks p . right     thread per gate column (coalesced rows of `right`)
//   stage 3  i,f,c,o pointwise + c/h update      thread per (sequence, unit)
// and the Dense top of the last layer.  Factor matrices are read through the read-only path (they are
// L1/L2 resident: <= a few MB for every configuration of BASELINE.json).
// Every activation array in shared memory is SEQUENCE-MINOR ([feature][BT]): the BT operands a thread multiplies one weight
// with are one or two 128-bit broadcast loads instead of BT scalar ones (the inner loops were LSU-bound: 1 + BT loads per BT
// FMAs).  The order of every floating-point sum is unchanged.
//
// This is the parity workhorse (every form x merged/split x mask/go_backwards/time_major/state);
// the latency-optimised batch-1 path is k1_wavefront.cu, the tensor-core path k1b_tc.cu.
// Replaces SingularLSTMCell.call / ReducedLSTMCell.call + the backend.rnn loop
// (reference code/svd_classes_v3.py:116-236, 317-368, 405-434).
#include "common.cuh"

namespace svdlstm {

namespace {

constexpr int kThreads = 256;

template <int BT>
__device__ __forceinline__ void store_bt(float* __restrict__ p, const float (&v)[BT]) {
  if constexpr (BT >= 4) {
#pragma unroll
    for (int q = 0; q < BT / 4; ++q) *reinterpret_cast<float4*>(p + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  } else {
#pragma unroll
    for (int q = 0; q < BT; ++q) p[q] = v[q];
  }
}

// acc[bt] += sum_k in[k][bt] * w[k * ld]  (k ascending, one fma per term: the order of the scalar loop).  The weights come from
// L2 (a layer's factors are far larger than L1): kUnroll loads are issued before the first is used, so each warp keeps
// kUnroll cache lines in flight instead of one or two -- the loop was bound by exactly that latency.
constexpr int kUnroll = 8;
template <int BT>
__device__ __forceinline__ void load_bt(const float* __restrict__ p, float (&v)[BT]);
template <int BT>
__device__ __forceinline__ void dot_bt(const float* __restrict__ w, int ld, const float* __restrict__ in, int n, float (&acc)[BT]) {
  int k = 0;
  for (; k + kUnroll <= n; k += kUnroll) {   // (requesting batch k+1 before consuming batch k was tried: slower, 52 -> 56 ms)
    float wv[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) wv[u] = __ldg(w + (size_t)(k + u) * ld);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float v[BT];
      load_bt<BT>(in + (k + u) * BT, v);
#pragma unroll
      for (int bt = 0; bt < BT; ++bt) acc[bt] = fmaf(v[bt], wv[u], acc[bt]);
    }
  }
  for (; k < n; ++k) {
    const float wk = __ldg(w + (size_t)k * ld);
    float v[BT];
    load_bt<BT>(in + k * BT, v);
#pragma unroll
    for (int bt = 0; bt < BT; ++bt) acc[bt] = fmaf(v[bt], wk, acc[bt]);
  }
}

// two weight columns against the same activations: half the shared-memory loads per FMA
template <int BT>
__device__ __forceinline__ void dot2_bt(const float* __restrict__ wa, const float* __restrict__ wb, int ld, const float* __restrict__ in, int n,
                                        float (&accA)[BT], float (&accB)[BT]) {
  constexpr int U = kUnroll;   // per column (4 deep per column: 47 -> 45 ms and much worse at small tiles; 16 deep: no further gain)
  int k = 0;
  for (; k + U <= n; k += U) {
    float va[U], vb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      va[u] = __ldg(wa + (size_t)(k + u) * ld);
      vb[u] = __ldg(wb + (size_t)(k + u) * ld);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float v[BT];
      load_bt<BT>(in + (k + u) * BT, v);
#pragma unroll
      for (int bt = 0; bt < BT; ++bt) {
        accA[bt] = fmaf(v[bt], va[u], accA[bt]);
        accB[bt] = fmaf(v[bt], vb[u], accB[bt]);
      }
    }
  }
  for (; k < n; ++k) {
    const float a = __ldg(wa + (size_t)k * ld), b = __ldg(wb + (size_t)k * ld);
    float v[BT];
    load_bt<BT>(in + k * BT, v);
#pragma unroll
    for (int bt = 0; bt < BT; ++bt) {
      accA[bt] = fmaf(v[bt], a, accA[bt]);
      accB[bt] = fmaf(v[bt], b, accB[bt]);
    }
  }
}

// BT consecutive floats (one feature of the CTA's BT sequences; 16-byte aligned when BT >= 4) -> registers
template <int BT>
__device__ __forceinline__ void load_bt(const float* __restrict__ p, float (&v)[BT]) {
  if constexpr (BT >= 4) {
#pragma unroll
    for (int q = 0; q < BT / 4; ++q) {
      const float4 t = *reinterpret_cast<const float4*>(p + 4 * q);
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
  } else if constexpr (BT == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1] = t.y;
  } else {
    v[0] = p[0];
  }
}

template <int BT, bool HAS_MASK>
__global__ void __launch_bounds__(kThreads) lstm_general_kernel(const ModelDesc* __restrict__ mdp, ForwardArgs a) {
  extern __shared__ __align__(16) float smem[];
  const ModelDesc& md = *mdp;
  const int L = md.n_layers;
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * BT;
  const int B = a.B, T = a.T;
  const bool ret_seq = a.flags & SVDLSTM_RETURN_SEQUENCES;
  const bool backwards = a.flags & SVDLSTM_GO_BACKWARDS;
  const bool time_major = a.flags & SVDLSTM_TIME_MAJOR;
  const bool zero_mask_out = a.flags & SVDLSTM_ZERO_OUTPUT_FOR_MASK;
  const int D = md.input_dim;

  // ---- shared-memory carve-up ---------------------------------------------------------------
  // (offsets, not pointers: `smem + offset` keeps the address space known to the compiler -> LDS/STS, not generic LD/ST)
  __shared__ int o_h[kMaxLayers];
  __shared__ int o_c[kMaxLayers];
  __shared__ int o_o[kMaxLayers];
#define s_h(l) (smem + o_h[l])
#define s_c(l) (smem + o_c[l])
#define s_o(l) (smem + o_o[l])
  int cur = 0;
  int maxp = 0;
  for (int l = 0; l < L; ++l) {
    const int H = md.layers[l].units;
    if (tid == 0) {
      o_h[l] = cur;
      o_c[l] = cur + BT * H;
      o_o[l] = HAS_MASK ? cur + 2 * BT * H : cur;
    }
    cur += (HAS_MASK ? 3 : 2) * BT * H;
    maxp = max(maxp, md.layers[l].p_total);
  }
  float* s_x = smem + cur;
  const int x_off = cur;
  cur += BT * D;
  float* s_p = smem + cur;
  __syncthreads();

  // ---- initial state ------------------------------------------------------------------------
  {
    size_t off = 0;
    for (int l = 0; l < L; ++l) {
      const int H = md.layers[l].units;
      for (int idx = tid; idx < BT * H; idx += kThreads) {
        const int bt = idx / H, j = idx - bt * H;
        const int b = b0 + bt;
        float hv = 0.f, cv = 0.f;
        if (a.h0 != nullptr && b < B) {
          hv = a.h0[off + (size_t)b * H + j];
          cv = a.c0[off + (size_t)b * H + j];
        }
        s_h(l)[j * BT + bt] = hv;
        s_c(l)[j * BT + bt] = cv;
        if (HAS_MASK) s_o(l)[j * BT + bt] = 0.f;
      }
      off += (size_t)B * H;
    }
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  const int n_y = md.n_out > 0 ? md.n_out : md.layers[L - 1].units;

  for (int step = 0; step < T; ++step) {
    const int t = backwards ? (T - 1 - step) : step;
    // x_t -> smem
    for (int idx = tid; idx < BT * D; idx += kThreads) {
      const int bt = idx / D, d = idx - bt * D;
      const int b = b0 + bt;
      float v = 0.f;
      if (b < B) v = time_major ? a.x[((size_t)t * B + b) * D + d] : a.x[((size_t)b * T + t) * D + d];
      s_x[d * BT + bt] = v;
    }
    __syncthreads();

    for (int l = 0; l < L; ++l) {
      const LayerDesc& Ld = md.layers[l];
      const int H = Ld.units;
      const int Din = Ld.d_in;
      const int lin_off = (l == 0) ? x_off : o_o[l - 1];
      const int h_off = o_h[l];
      const int P = Ld.p_total;

      // ---- stage 1 ---------------------------------------------------------------------------
      for (int q = tid; q < P; q += kThreads) {
        int bi = 0;
        while (bi + 1 < Ld.n_blocks && q >= Ld.blocks[bi + 1].p_off) ++bi;
        const Block& blk = Ld.blocks[bi];
        const int k = q - blk.p_off;
        const float* in = smem + (blk.from_h ? h_off : lin_off);
        const int kin = blk.from_h ? H : Din;
        float acc[BT];
        if (blk.left == nullptr) {
#pragma unroll
          for (int bt = 0; bt < BT; ++bt) acc[bt] = in[k * BT + bt];
        } else {
#pragma unroll
          for (int bt = 0; bt < BT; ++bt) acc[bt] = 0.f;
          dot_bt<BT>(blk.left + k, blk.left_ld, in, kin, acc);
          if (blk.scale != nullptr) {
            const float s = __ldg(blk.scale + k);
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) acc[bt] *= s;
          }
        }
        store_bt<BT>(s_p + q * BT, acc);
      }
      __syncthreads();

      // ---- stages 2 + 3: one thread = the four gate columns of one unit, for the CTA's BT sequences; gates + state update
      //      (Keras _compute_carry_and_output_fused) straight from the accumulators. ----------------------------------------------
      if (4 * H <= kThreads) {
        // Narrow layers (the shipped model: 15 units): a thread per UNIT would leave most of the CTA idle behind H long serial
        // chains -- a thread per gate COLUMN instead (quad of lanes = the i, f, c~, o columns of one unit), the quad's four
        // pre-activations gathered by shuffles, lane 0 of the quad updating the cell.
        const int j = tid >> 2, q = tid & 3;
        const bool act = j < H;
        float acc[BT];
#pragma unroll
        for (int bt = 0; bt < BT; ++bt) acc[bt] = 0.f;
        if (act) {
          const int n = q * H + j;
          const float bv = __ldg(Ld.bias + n);
#pragma unroll
          for (int bt = 0; bt < BT; ++bt) acc[bt] = bv;
          for (int bi = 0; bi < Ld.n_blocks; ++bi) {
            const Block& blk = Ld.blocks[bi];
            int rel = n - blk.out0;
            if (rel < 0) continue;
            const float* pp = s_p + blk.p_off * BT;
            if (blk.ident) {
              if (rel < blk.rank) {
#pragma unroll
                for (int bt = 0; bt < BT; ++bt) acc[bt] += pp[rel * BT + bt];
                continue;
              }
              rel -= blk.rank;
            }
            if (rel >= blk.ncols) continue;
            dot_bt<BT>(blk.right + rel, blk.right_ld, pp, blk.rank, acc);
          }
        }
        const int quad0 = (tid & 31) & ~3;
#pragma unroll
        for (int bt = 0; bt < BT; ++bt) {
          const float zi = __shfl_sync(0xffffffffu, acc[bt], quad0), zf = __shfl_sync(0xffffffffu, acc[bt], quad0 + 1);
          const float zc = __shfl_sync(0xffffffffu, acc[bt], quad0 + 2), zo = __shfl_sync(0xffffffffu, acc[bt], quad0 + 3);
          if (act && q == 0) {
            const int idx = j * BT + bt;
            const float c_new = sigmoid_acc(zf) * s_c(l)[idx] + sigmoid_acc(zi) * tanhf(zc);
            const float hv = sigmoid_acc(zo) * tanhf(c_new);
            if (HAS_MASK) {
              const int b = b0 + bt;
              const bool valid = (b < B) ? (a.mask[(size_t)b * T + t] != 0) : true;
              if (valid) {
                s_c(l)[idx] = c_new;
                s_h(l)[idx] = hv;
                s_o(l)[idx] = hv;
              } else if (zero_mask_out) {
                s_o(l)[idx] = 0.f;
              }
            } else {
              s_c(l)[idx] = c_new;
              s_h(l)[idx] = hv;
            }
          }
        }
      } else
      for (int j = tid; j < H; j += kThreads) {
        float keep[BT];   // sigmoid(i) -> sigmoid(i) tanh(c~) -> c_new
        float h_new[BT];
#pragma unroll
        for (int gp = 0; gp < 2; ++gp) {   // gate columns in pairs (i, c~) then (f, o): one load of p serves both
          const int nA = (gp == 0 ? 0 : 1) * H + j, nB = (gp == 0 ? 2 : 3) * H + j;
          float accA[BT], accB[BT];
          const float bA = __ldg(Ld.bias + nA), bB = __ldg(Ld.bias + nB);
#pragma unroll
          for (int bt = 0; bt < BT; ++bt) { accA[bt] = bA; accB[bt] = bB; }
          for (int bi = 0; bi < Ld.n_blocks; ++bi) {
            const Block& blk = Ld.blocks[bi];
            const float* pp = s_p + blk.p_off * BT;
            // where a column falls in this block: 0 nowhere, 1 the identity part of a 2-factor block, 2 the right factor
            int relA = nA - blk.out0, relB = nB - blk.out0;
            int kindA = relA < 0 ? 0 : 2, kindB = relB < 0 ? 0 : 2;
            if (blk.ident) {
              if (kindA) { if (relA < blk.rank) kindA = 1; else relA -= blk.rank; }
              if (kindB) { if (relB < blk.rank) kindB = 1; else relB -= blk.rank; }
            }
            if (kindA == 2 && relA >= blk.ncols) kindA = 0;
            if (kindB == 2 && relB >= blk.ncols) kindB = 0;
            if (kindA == 2 && kindB == 2) {
              dot2_bt<BT>(blk.right + relA, blk.right + relB, blk.right_ld, pp, blk.rank, accA, accB);
              continue;
            }
            if (kindA == 1) {
#pragma unroll
              for (int bt = 0; bt < BT; ++bt) accA[bt] += pp[relA * BT + bt];
            } else if (kindA == 2) {
              dot_bt<BT>(blk.right + relA, blk.right_ld, pp, blk.rank, accA);
            }
            if (kindB == 1) {
#pragma unroll
              for (int bt = 0; bt < BT; ++bt) accB[bt] += pp[relB * BT + bt];
            } else if (kindB == 2) {
              dot_bt<BT>(blk.right + relB, blk.right_ld, pp, blk.rank, accB);
            }
          }
          const float* c_old = s_c(l) + j * BT;
#pragma unroll
          for (int bt = 0; bt < BT; ++bt) {
            if (gp == 0) keep[bt] = sigmoid_acc(accA[bt]) * tanhf(accB[bt]);
            else {
              keep[bt] = sigmoid_acc(accA[bt]) * c_old[bt] + keep[bt];
              h_new[bt] = sigmoid_acc(accB[bt]) * tanhf(keep[bt]);
            }
          }
        }
#pragma unroll
        for (int bt = 0; bt < BT; ++bt) {
          const int idx = j * BT + bt;
          if (HAS_MASK) {
            const int b = b0 + bt;
            const bool valid = (b < B) ? (a.mask[(size_t)b * T + t] != 0) : true;
            if (valid) {
              s_c(l)[idx] = keep[bt];
              s_h(l)[idx] = h_new[bt];
              s_o(l)[idx] = h_new[bt];
            } else if (zero_mask_out) {
              s_o(l)[idx] = 0.f;
            }
          } else {
            s_c(l)[idx] = keep[bt];
            s_h(l)[idx] = h_new[bt];
          }
        }
      }
      __syncthreads();
    }

    // ---- output of this step -----------------------------------------------------------------
    if (ret_seq || step == T - 1) {
      const float* ho = s_o(L - 1);
      const int HL = md.layers[L - 1].units;
      if (md.n_out > 0) {
        for (int pair = warp; pair < BT * md.n_out; pair += kThreads / 32) {
          const int bt = pair / md.n_out, o = pair - bt * md.n_out;
          float acc = 0.f;
          for (int j = lane; j < HL; j += 32) acc = fmaf(ho[j * BT + bt], __ldg(md.dense_kernel + (size_t)j * md.n_out + o), acc);
#pragma unroll
          for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
          const int b = b0 + bt;
          if (lane == 0 && b < B) {
            acc += __ldg(md.dense_bias + o);
            size_t yi;
            if (!ret_seq) yi = (size_t)b * n_y + o;
            else yi = time_major ? ((size_t)step * B + b) * n_y + o : ((size_t)b * T + step) * n_y + o;
            a.y[yi] = acc;
          }
        }
      } else {
        for (int idx = tid; idx < BT * HL; idx += kThreads) {
          const int bt = idx / HL, j = idx - bt * HL;
          const int b = b0 + bt;
          if (b < B) {
            size_t yi;
            if (!ret_seq) yi = (size_t)b * n_y + j;
            else yi = time_major ? ((size_t)step * B + b) * n_y + j : ((size_t)b * T + step) * n_y + j;
            a.y[yi] = ho[j * BT + bt];
          }
        }
      }
    }
    // no barrier needed here: the next step's first write (s_x) is separated from this step's
    // reads of s_o[L-1] by the barrier after the s_x load only if L>1 ... keep it simple and safe:
    __syncthreads();
  }

  // ---- final state ---------------------------------------------------------------------------
  if (a.h_n != nullptr || a.c_n != nullptr) {
    size_t off = 0;
    for (int l = 0; l < L; ++l) {
      const int H = md.layers[l].units;
      for (int idx = tid; idx < BT * H; idx += kThreads) {
        const int bt = idx / H, j = idx - bt * H;
        const int b = b0 + bt;
        if (b < B) {
          if (a.h_n) a.h_n[off + (size_t)b * H + j] = s_h(l)[j * BT + bt];
          if (a.c_n) a.c_n[off + (size_t)b * H + j] = s_c(l)[j * BT + bt];
        }
      }
      off += (size_t)B * H;
    }
  }
}

#undef s_h
#undef s_c
#undef s_o

size_t general_smem_bytes(const ModelDesc& md, int BT, bool has_mask) {
  size_t f = 0;
  int maxp = 0;
  for (int l = 0; l < md.n_layers; ++l) {
    const int H = md.layers[l].units;
    f += (size_t)(has_mask ? 3 : 2) * BT * H;
    if (md.layers[l].p_total > maxp) maxp = md.layers[l].p_total;
  }
  f += (size_t)BT * md.input_dim + (size_t)BT * maxp;
  return f * sizeof(float);
}

template <int BT>
int launch_general(const ModelDesc& md, const ModelDesc* dev_md, const ForwardArgs& a, size_t smem, cudaStream_t stream) {
  const int grid = (a.B + BT - 1) / BT;
  if (a.mask) {
    SVD_CUDA_TRY(cudaFuncSetAttribute(lstm_general_kernel<BT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_general_kernel<BT, true><<<grid, kThreads, smem, stream>>>(dev_md, a);
  } else {
    SVD_CUDA_TRY(cudaFuncSetAttribute(lstm_general_kernel<BT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_general_kernel<BT, false><<<grid, kThreads, smem, stream>>>(dev_md, a);
  }
  SVD_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace

int run_general(const ModelDesc& md, const ModelDesc* dev_md, const ForwardArgs& a, cudaStream_t stream, int* launches) {
  const size_t kMaxSmem = 227 * 1024;
  // sequences per CTA: every weight load (L2) serves BT sequences, so as many as still leave ~one CTA per SM, subject to shared
  // memory (measured, C3 rank 128, T=256: B=4096 BT 4/8/16 = 73/58/52 ms; B=512 BT 1/4/8/16 = 38/26/28/33 ms)
  int BT = 16;
  while (BT > 1 && ((a.B + BT - 1) / BT < 120 || general_smem_bytes(md, BT, a.mask != nullptr) > kMaxSmem)) BT >>= 1;
  if (const char* e = getenv("SVDLSTM_GEN_BT")) {   // experiments
    const int v = atoi(e);
    if ((v == 1 || v == 2 || v == 4 || v == 8 || v == 16) && general_smem_bytes(md, v, a.mask != nullptr) <= kMaxSmem) BT = v;
  }
  const size_t smem = general_smem_bytes(md, BT, a.mask != nullptr);
  SVD_REQUIRE(smem <= kMaxSmem, "svdlstm_forward(general): model state needs %zu B of shared memory per sequence (> %zu)", smem, kMaxSmem);
  int rc;
  switch (BT) {
    case 16: rc = launch_general<16>(md, dev_md, a, smem, stream); break;
    case 8: rc = launch_general<8>(md, dev_md, a, smem, stream); break;
    case 4: rc = launch_general<4>(md, dev_md, a, smem, stream); break;
    case 2: rc = launch_general<2>(md, dev_md, a, smem, stream); break;
    default: rc = launch_general<1>(md, dev_md, a, smem, stream); break;
  }
  if (rc == 0) *launches = 1;
  return rc;
}

}  // namespace svdlstm
