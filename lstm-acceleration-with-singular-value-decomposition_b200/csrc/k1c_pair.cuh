// K1c: the paired-CTA (tcgen05 cta_group::2) form of the tensor-core recurrent kernel.  Included by k1b_tc.cu
// (same translation unit: it shares the PTX wrappers, operand-image helpers and the per-device workspace).
//
// Why: with weights re-streamed from L2 every step (ranks >= 64 at H = 256) the single-CTA kernel is bound by the
// shared-memory port -- every 16 KB weight chunk is written once and read once for only 64 sequences (DESIGN.md
// section 4.3).  Here TWO CTAs on the two SMs of one TPC own ONE tile of 128 sequences of one layer:
//   * every MMA is a cta_group::2 instruction, M = 256, N = 128: each CTA supplies 128 weight rows (A) and the
//     activations of its 64 sequences (B half); accumulators for its 128 rows x all 128 sequences land in its TMEM.
//     Each SM therefore streams only HALF of the layer's weights, and every streamed byte serves 128 sequences.
//   * CTA c owns cells [c H/2, (c+1) H/2): its S2 tiles are the i, g, f, o gate rows of its own cells, so the cell
//     update needs no exchange; h(t) of own cells x 128 sequences is written straight into the B-operand images:
//     columns 0..63 into CTA 0's shared memory, 64..127 into CTA 1's (st.shared::cluster).
//   * S1 is ONE M = 256 tile over K = H with BOTH halves useful: rows 0..127 (CTA 0) = t_u of this layer for the
//     next step, rows 128..255 (CTA 1) = "X rows" computed from the same h(t-1): the NEXT layer's t_w(t-1)
//     (handed over through HBM already in B-operand layout -- half the bytes of handing h over) or, on the last
//     layer, the Dense-top rows y(t-1).  Ranks above 128 use two such tiles.
//   * S2 accumulators rotate through two 128-column TMEM slots: i -> A, g -> B, f -> A, o -> B.
// Only the leader CTA (cluster rank 0) issues MMAs; the peer's warp 1 forwards "my ring slot / input tile landed"
// to the leader's barriers.  Replaces the same reference lines as k1b_tc.cu.

constexpr int kPairEpiWarps = 16;
constexpr int kPairThreads = 32 * (4 + kPairEpiWarps);   // 640
constexpr int kPairMaxSlots = 12;
constexpr int kPairInStagesMax = 3;
constexpr uint32_t kPairSlotBytes = 16384;   // one weight chunk (<= 4 MMAs) per ring slot

// ------------------------------------------------------------------------------------------------
// cluster PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on a barrier given by its shared::cluster address (own or peer CTA), cluster-scope release
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cbar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cbar) : "memory");
}
// arrive that publishes no generic-proxy data of its own: "the bulk copy into my shared memory completed" (the forwarder saw the
// phase flip, i.e. the bytes had landed before this arrive was even issued) or "my tcgen05.ld of this TMEM slot completed" (ordered by
// tcgen05.fence::before_thread_sync).  The release form costs MEMBAR.ALL.GPU + error barriers: ~1 us per arrive on a single thread.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cbar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cbar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cl(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded, abortable waits.  A protocol bug must neither hang the GPU nor kill the context before it can be diagnosed: the
// first wait that times out records (site, block, thread, parity) in g_pair_dbg and raises g_pair_abort, after which every
// wait of the launch returns at once (the results are garbage; run_tc_pair reports the record as an error).
__device__ unsigned int g_pair_abort;
__device__ unsigned int g_pair_nrec;
__device__ unsigned int g_pair_dbg[4 * 32];   // up to 32 records: (site | step << 8, block, thread, parity)
__device__ __noinline__ bool pair_wait_slow(uint32_t bar, uint32_t parity, uint32_t site, bool cl) {
  const long long t0 = clock64();
  uint32_t spins = 0;
  bool recorded = false;
  while (!(cl ? mbar_try_wait_cl(bar, parity) : mbar_try_wait(bar, parity))) {
    if ((++spins & 0x3FFu) == 0) {   // the hot spin is try_wait only: a global load per spin would add its latency to every wake-up
      const long long dt = clock64() - t0;
      if (dt > 100000000LL && *reinterpret_cast<volatile unsigned int*>(&g_pair_abort)) return true;
      if (dt > 1000000000LL && !recorded) {   // every stuck waiter leaves a record ...
        recorded = true;
        const unsigned int i = atomicAdd(&g_pair_nrec, 1u);
        if (i < 32u) {
          g_pair_dbg[4 * i + 0] = site;
          g_pair_dbg[4 * i + 1] = blockIdx.x;
          g_pair_dbg[4 * i + 2] = threadIdx.x;
          g_pair_dbg[4 * i + 3] = parity;
        }
        __threadfence();
      }
      if (dt > 1400000000LL) {                // ... before the launch is abandoned
        atomicExch(&g_pair_abort, 1u);
        return true;
      }
    }
  }
  return false;
}
// `dead` (one per thread) becomes true once a wait was abandoned; every role loop leaves at its next iteration.
#define pwait(bar_, parity_, site_)    do { if (!mbar_try_wait(bar_, parity_)) dead |= pair_wait_slow(bar_, parity_, site_, false); } while (0)
#define pwait_cl(bar_, parity_, site_) do { if (!mbar_try_wait_cl(bar_, parity_)) dead |= pair_wait_slow(bar_, parity_, site_, true); } while (0)
// 16-byte store into the shared memory of a CTA of the cluster (own or peer) whose completion is counted (complete_tx, 16 bytes) on
// an mbarrier of the SAME destination CTA: the signal travels with the data, so the writers need no fence and no arrive -- a
// release.cluster arrive costs MEMBAR.ALL.GPU, ~0.5 us per warp and step on the recurrence's critical path.
__device__ __forceinline__ void sta128_cluster(uint32_t caddr, uint32_t cbar, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(caddr), "r"(a), "r"(b),
               "r"(c), "r"(d), "r"(cbar)
               : "memory");
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit of all prior MMAs of this thread -> one arrival on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit2(uint32_t bar, uint32_t elected) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      ".reg .b16 m;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "mov.b16 m, 3;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t"
      "}\n" ::"r"(bar),
      "r"(elected)
      : "memory");
}
// NM (1..4) back-to-back pair MMAs (K = 16 each) of one weight chunk: A +256 B per K step (lo +16), B +2 k-groups of a
// 64-sequence half tile = 2048 B (lo +128).
#define SVD_UMMA2_HEAD                                                  \
  "{\n\t"                                                               \
  ".reg .pred p, q, one;\n\t"                                           \
  ".reg .b64 da, db;\n\t"                                               \
  ".reg .b32 ta, tb;\n\t"                                               \
  "setp.ne.b32 q, %7, 0;\n\t"                                           \
  "setp.ne.b32 p, %6, 0;\n\t"                                           \
  "setp.eq.b32 one, 0, 0;\n\t"                                          \
  "mov.b64 da, {%1, %2};\n\t"                                           \
  "mov.b64 db, {%3, %4};\n\t"                                           \
  "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
#define SVD_UMMA2_NEXT(aoff, boff)                                      \
  "add.u32 ta, %1, " #aoff ";\n\t"                                      \
  "add.u32 tb, %3, " #boff ";\n\t"                                      \
  "mov.b64 da, {ta, %2};\n\t"                                           \
  "mov.b64 db, {tb, %4};\n\t"                                           \
  "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, one;\n\t"
template <int NM>
__device__ __forceinline__ void umma2_f16_x(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                            uint32_t first_accumulates, uint32_t elected) {
  static_assert(NM >= 1 && NM <= 4, "1..4 MMAs per chunk");
  if constexpr (NM == 1) asm volatile(SVD_UMMA2_HEAD "}\n" SVD_UMMA_OPERANDS);
  if constexpr (NM == 2) asm volatile(SVD_UMMA2_HEAD SVD_UMMA2_NEXT(16, 128) "}\n" SVD_UMMA_OPERANDS);
  if constexpr (NM == 3) asm volatile(SVD_UMMA2_HEAD SVD_UMMA2_NEXT(16, 128) SVD_UMMA2_NEXT(32, 256) "}\n" SVD_UMMA_OPERANDS);
  if constexpr (NM == 4) asm volatile(SVD_UMMA2_HEAD SVD_UMMA2_NEXT(16, 128) SVD_UMMA2_NEXT(32, 256) SVD_UMMA2_NEXT(48, 384) "}\n" SVD_UMMA_OPERANDS);
}

// ------------------------------------------------------------------------------------------------
// parameters
// ------------------------------------------------------------------------------------------------
enum { PK_NONE = 0, PK_U = 1, PK_X = 2, PK_DENSE = 3, PK_S2_LATE = 4, PK_S2_EARLY = 5, PK_S2_EARLY0 = 6 };

struct PairChunk {      // packer work item: one weight chunk of one CTA rank ([128 rows x kc], K-major)
  uint32_t byte_off;
  int16_t kind, rows, kc, r0, k0, gate, cell0, pad;
};

struct PairLayerParams {
  const uint8_t* wimg;       // weight stream image of both ranks
  const uint32_t* slots;     // [2][n_s1 + n_s2] ring-slot fills per step: (offset >> 8) | (bytes >> 8) << 16
  int n_s1, n_s2;
  const float* bias;         // [H/128][4][128], slots i, g, f, o; sigmoid gates pre-scaled by 0.5
  const uint8_t* in_seq;     // input tiles [2 tile + rank][t], K = kin (layer 0: x image; else the t_w image of the previous layer)
  uint8_t* out_seq;          // next layer's t_w tiles [2 tile + half][t], K = x_pad                 (non-last layers)
  float* y;                  // (B, T, n_dense)                                                      (last layer)
  const float* dense_bias;
  int* prog_in;              // [tile] contributions published by the previous layer (nullptr: input complete)
  int* prog_out;             // [tile]
  int in_contrib;            // contributions per step of the producer (1 or 2)
  int H, T, B;
  int kin;                   // K extent of the early (input) part of S2, multiple of 16
  int ru, ru_pad;
  int n_mt;                  // S1 M = 256 tiles
  int half_kind[2][2];       // [mt][rank]
  int half_r0[2][2];
  int x_rows, x_pad;         // rows of the next layer's t_w (0 on the last layer), padded to 16
  int n_dense;
  int in_stages, w_slots;
};

struct PairSmemPlan {
  uint32_t w, hbuf, tbuf, inbuf, xstage, bars, tmem_slot, ctab, total;
};

__host__ __device__ inline PairSmemPlan pair_plan(const PairLayerParams& p) {
  PairSmemPlan s;
  uint32_t off = 0;
  s.w = off; off += (uint32_t)p.w_slots * kPairSlotBytes;
  s.hbuf = off; off += act_tile_bytes(p.H, 64);
  s.tbuf = off; off += act_tile_bytes(p.ru_pad, 64);
  s.inbuf = off; off += (uint32_t)p.in_stages * act_tile_bytes(p.kin, 64);
  s.xstage = off; off += p.x_rows > 0 ? 2u * act_tile_bytes(128, 64) : 0u;
  s.bars = off; off += 512;
  s.tmem_slot = off; off += 16;
  s.ctab = off; off += (uint32_t)(((p.n_s1 + p.n_s2) * 4 + 15) & ~15);
  s.total = off;
  return s;
}

enum {
  PB_W_FULL = 0,
  PB_W_EMPTY = kPairMaxSlots,
  PB_IN_FULL = 2 * kPairMaxSlots,
  PB_IN_EMPTY = PB_IN_FULL + kPairInStagesMax,
  PB_S1_FULL = PB_IN_EMPTY + kPairInStagesMax,
  PB_T_READY,
  PB_S2_FULL0,
  PB_S2_FULL1,
  PB_S2_EMPTY0,
  PB_S2_EMPTY1,
  PB_H_READY,
  PB_X_STAGED,
  PB_X_STORED,
  PB_COUNT
};
static_assert(PB_COUNT * 8 <= 512, "barrier block too small");

struct PairPipeParams {
  PairLayerParams layer[kMaxLayers];
  int n_layers, n_tiles;
  long long* tl;   // debug build (SVDLSTM_TC_TIMELINE): globaltimer stamps of step 40, tile 0, [layer][64]
  int off;     // bring-up: role mask switched off (1 streamer, 2 input loader, 4 X store, 8 MMA issuer, 16 forwarder, 32 epilogue)
  int stage;   // bring-up bisection (SVDLSTM_PAIR_STAGE): 0 = everything; 1 = setup/teardown only; 2 = barrier protocol only (no MMAs, no TMEM
               // loads, no operand stores); 3 = + MMAs; 4 = + TMEM loads; 5 = + stores (= 0)
};

// S2 chunk order of one step, shared by the packer (host) and the MMA warp: per own unit block, per gate pair
// (i, g) then (f, o): [slot-0 tile early][slot-1 tile early][slot-0 tile late][slot-1 tile late], chunks of <= 64 K.
// f(gate_slot 0..3, ub, late, k0, kc)
template <class F>
__host__ __device__ inline void for_pair_s2(int nubc, int kin, int ru_pad, F&& f) {
  for (int ub = 0; ub < nubc; ++ub)
    for (int gp = 0; gp < 2; ++gp) {
      for (int sl = 0; sl < 2; ++sl)
        for (int k0 = 0; k0 < kin; k0 += 64) f(gp * 2 + sl, ub, 0, k0, imin(64, kin - k0));
      for (int sl = 0; sl < 2; ++sl)
        for (int k0 = 0; k0 < ru_pad; k0 += 64) f(gp * 2 + sl, ub, 1, k0, imin(64, ru_pad - k0));
    }
}

#ifdef SVDLSTM_TC_TIMELINE
__device__ __forceinline__ long long gtime_ns() {
  long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}
#define PT_STAMP(slot) do { if (pp.tl != nullptr && tile == 0 && tdbg == 40) pp.tl[layer * 64 + (slot)] = gtime_ns(); } while (0)
#define PT_ADD(slot, v) do { if (pp.tl != nullptr && tile == 0 && tdbg == 40) pp.tl[layer * 64 + (slot)] += (v); } while (0)
#define PT_NOW() gtime_ns()
#else
#define PT_STAMP(slot) do { } while (0)
#define PT_ADD(slot, v) do { } while (0)
#define PT_NOW() 0ll
#endif

// ------------------------------------------------------------------------------------------------
// the kernel: CTA pair = one layer x one tile of 128 sequences x all T steps
// ------------------------------------------------------------------------------------------------
template <int NUBC>   // unit blocks of 128 cells owned by each CTA (H = 256 NUBC)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1) lstm_tc_pair_kernel(const __grid_constant__ PairPipeParams pp) {
  constexpr uint32_t kK64 = 64u * 8u;             // descriptor-lo step of 64 K rows of a 64-sequence half tile (8 k-groups x 1024 B)
  constexpr uint32_t kRowBlk = 64u * 256u;        // bytes of 128 K rows of a 64-sequence half tile
  extern __shared__ __align__(1024) uint8_t smem[];
  const int pair = (int)blockIdx.x >> 1;
  const int layer = pair / pp.n_tiles, tile = pair - layer * pp.n_tiles;
  const PairLayerParams& p = pp.layer[layer];
  const uint32_t rank = cluster_ctarank();
  const PairSmemPlan sp = pair_plan(p);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, T = p.T;
  const int nst = p.in_stages, ns = p.w_slots;
  const uint32_t in_tile = act_tile_bytes(p.kin, 64);
  const uint32_t bar0 = sbase + sp.bars;
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int n_steps = T + 1;   // the X rows (next layer's t_w / Dense top) of step T-1 fall out of one more S1 pass
  const int n_fill_step = p.n_s1 + p.n_s2;
  const int* prog_in = p.prog_in ? p.prog_in + tile : nullptr;
  int* prog_out = p.prog_out ? p.prog_out + tile : nullptr;
  int tdbg = 0;   // step index recorded by a timed-out wait
  bool dead = false;

  // ---- one-time setup ---------------------------------------------------------------------------
  for (uint32_t i = threadIdx.x * 16u; i < sp.bars; i += kPairThreads * 16u)   // ring + activation buffers: zeros (h(-1) = 0; everything finite)
    *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  {
    uint32_t* ctab = reinterpret_cast<uint32_t*>(smem + sp.ctab);
    for (int i = threadIdx.x; i < n_fill_step; i += kPairThreads) ctab[i] = p.slots[(size_t)rank * n_fill_step + i];
  }
  if (threadIdx.x == 0) {
    const uint32_t both = rank == 0 ? 2u : 1u;    // leader: own producer + the peer's forwarder
    for (int s = 0; s < kPairMaxSlots; ++s) {
      mbar_init(bar(PB_W_FULL + s), both);
      mbar_init(bar(PB_W_EMPTY + s), 1);
    }
    for (int s = 0; s < kPairInStagesMax; ++s) {
      mbar_init(bar(PB_IN_FULL + s), both);
      mbar_init(bar(PB_IN_EMPTY + s), 1);
    }
    mbar_init(bar(PB_S1_FULL), 1);
    mbar_init(bar(PB_T_READY), both);   // tx-counted operand halves: own tracker's expect_tx (+ on the leader: the peer tracker's "my half landed")
    mbar_init(bar(PB_S2_FULL0), 1);
    mbar_init(bar(PB_S2_FULL1), 1);
    mbar_init(bar(PB_S2_EMPTY0), 2 * kPairEpiWarps);
    mbar_init(bar(PB_S2_EMPTY1), 2 * kPairEpiWarps);
    mbar_init(bar(PB_H_READY), both);
    mbar_init(bar(PB_X_STAGED), kPairEpiWarps);
    mbar_init(bar(PB_X_STORED), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(sbase + sp.tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs: barriers initialised, buffers zeroed, TMEM allocated
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + sp.tmem_slot);
  // TMEM columns: [0,128) S1 tile 0, [128,256) S1 tile 1, [256,384) S2 slot 0 (gates i, f), [384,512) S2 slot 1 (gates g, o)
  const uint32_t tm_s1 = tmem, tm_s2 = tmem + 256u;

  const int stage = pp.stage;
  const bool ep_ld = stage == 0 || stage >= 4, ep_st = stage == 0 || stage >= 5;
  if (stage == 1) {
    // bring-up: nothing but setup and teardown
  } else if ((warp == 0 && (pp.off & 1)) || (warp == 3 && (pp.off & 2)) || (warp == 2 && (pp.off & 4)) || (warp == 1 && rank == 0 && (pp.off & 8)) ||
             (warp == 1 && rank == 1 && (pp.off & 16)) || (warp >= 4 && (pp.off & 32))) {
    // bring-up: role switched off
  } else if (warp == 0) {
    // ======================= weight streamer (each CTA streams its own half of the rows) ===========
    if (lane == 0) {
      const uint8_t* wimg = p.wimg;
      int slot = 0;
      uint32_t use = 0;
#pragma unroll 1
      for (int t = 0; t < n_steps; ++t) {
        tdbg = t;
        if (dead) break;
        const int n = t < T ? n_fill_step : p.n_s1;
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
          if (dead) break;
          const uint32_t e = reinterpret_cast<const volatile uint32_t*>(smem + sp.ctab)[i];
          const uint32_t bytes = (e >> 16) << 8;
          if (use > 0) pwait(bar(PB_W_EMPTY + slot), (use - 1u) & 1u, 1u | ((uint32_t)tdbg << 8));
          if (bytes > 0) {
            mbar_expect_tx(bar(PB_W_FULL + slot), bytes);
            bulk_g2s(sbase + sp.w + (uint32_t)slot * kPairSlotBytes, wimg + ((size_t)(e & 0xFFFFu) << 8), bytes, bar(PB_W_FULL + slot));
          } else {
            mbar_arrive(bar(PB_W_FULL + slot));   // this half has no rows in the chunk: the slot keeps its (finite) stale bytes
          }
          if (++slot == ns) { slot = 0; ++use; }
        }
      }
    }
  } else if (warp == 3) {
    // ======================= input ring: this CTA's 64-sequence half tile of step t ====================
    if (lane == 0) {
      const uint8_t* src = p.in_seq + (size_t)(2 * tile + (int)rank) * T * in_tile;
      int ld_s = 0;
      uint32_t ld_n = 0;
      const uint32_t lead_in = mapa_u32(bar(PB_IN_FULL), 0);
      // peer CTA: tell the leader's barrier that tile k landed in MY shared memory (the thread that issued the copy does it: a second
      // blocked lane in the forwarder warp would starve the weight-slot forwarding)
      auto forward_in = [&](int k) {
        const int st_k = k % nst;
        pwait(bar(PB_IN_FULL + st_k), (uint32_t)(k / nst) & 1u, 10u | ((uint32_t)tdbg << 8));
        if (!dead) mbar_arrive_cluster_relaxed(lead_in + 8u * (uint32_t)st_k);
      };
#pragma unroll 1
      for (int ld_t = 0; ld_t < T; ++ld_t) {
        tdbg = ld_t;
        if (dead) break;
        if (rank == 1 && ld_t > 0) forward_in(ld_t - 1);
        if (prog_in != nullptr) {
          const int need = p.in_contrib * (ld_t + 1);
          if (ld_acquire_gpu(prog_in) < need) {
            const long long t0 = clock64();
            while (ld_acquire_gpu(prog_in) < need) {
              if (*reinterpret_cast<volatile unsigned int*>(&g_pair_abort)) { dead = true; break; }
              if (clock64() - t0 > 2000000000LL) {
                const unsigned int i = atomicAdd(&g_pair_nrec, 1u);
                if (i < 32u) {
                  g_pair_dbg[4 * i + 0] = 99u | ((uint32_t)ld_t << 8);
                  g_pair_dbg[4 * i + 1] = blockIdx.x;
                  g_pair_dbg[4 * i + 2] = threadIdx.x;
                  g_pair_dbg[4 * i + 3] = (unsigned)ld_acquire_gpu(prog_in);
                }
                atomicExch(&g_pair_abort, 1u);
                dead = true;
                break;
              }
            }
          }
          fence_proxy_async_all();
          if (dead) break;
        }
        if (ld_n > 0) pwait(bar(PB_IN_EMPTY + ld_s), (ld_n - 1u) & 1u, 2u | ((uint32_t)tdbg << 8));
        mbar_expect_tx(bar(PB_IN_FULL + ld_s), in_tile);
        bulk_g2s(sbase + sp.inbuf + ld_s * in_tile, src + (size_t)ld_t * in_tile, in_tile, bar(PB_IN_FULL + ld_s));
        if (++ld_s == nst) { ld_s = 0; ++ld_n; }
      }
      if (rank == 1 && T > 0 && !dead) forward_in(T - 1);
    }
  } else if (warp == 2) {
    // ======================= X-row stores: this CTA's rows of the next layer's t_w(t), both sequence halves ==========
    int xr0 = -1;
    for (int mt = 0; mt < p.n_mt; ++mt)
      if (p.half_kind[mt][rank] == PK_X) xr0 = p.half_r0[mt][rank];
    if (lane == 0 && xr0 >= 0) {
      const uint32_t x_tile = act_tile_bytes(p.x_pad, 64);
      const uint32_t own_bytes = (uint32_t)imin(128, p.x_pad - xr0) * 128u;   // rows x 64 sequences x 2 B, contiguous in the tile image
      const uint32_t row_off = (uint32_t)(xr0 >> 3) * 1024u;
#pragma unroll 1
      for (int t = 0; t < T; ++t) {
        tdbg = t;
        pwait(bar(PB_X_STAGED), (uint32_t)t & 1u, 3u | ((uint32_t)tdbg << 8));
        if (dead) break;
#pragma unroll
        for (int c = 0; c < 2; ++c)
          bulk_s2g(p.out_seq + ((size_t)(2 * tile + c) * T + t) * x_tile + row_off, sbase + sp.xstage + (uint32_t)c * kRowBlk, own_bytes);
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(bar(PB_X_STORED));
        if (prog_out != nullptr) {
          bulk_wait_all0();
          fence_proxy_async_all();
          red_release_add(prog_out, 1);
        }
      }
      bulk_wait_all0();
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ======================= MMA issuer (leader CTA; whole warp converged, one elected lane issues) =============
      const uint32_t elected = elect_one();
      const uint32_t mma_el = (stage == 0 || stage >= 3) ? elected : 0u;
      const uint32_t idesc = make_idesc(256, 128);
      const uint32_t act_hi = desc_hi(kActSBO);
      constexpr uint32_t kActLBO = 64u * 16u;
      const uint32_t h_lo0 = desc_lo(sbase + sp.hbuf, kActLBO), t_lo0 = desc_lo(sbase + sp.tbuf, kActLBO);
      const uint32_t in_lo0 = desc_lo(sbase + sp.inbuf, kActLBO), in_stage_lo = in_tile >> 4;
      const uint32_t w_lo0 = desc_lo(sbase + sp.w, 128u);
      const uint32_t a_hi64 = desc_hi(1024u);
      int w_slot = -1;
      uint32_t w_use = 0;
      auto next_slot = [&]() -> uint32_t {
        if (w_slot >= 0) {
          if (pp.off & 1024) {   // timing experiment (stage 2 only): plain arrives instead of the multicast commit
            if (elected) {
              mbar_arrive(bar(PB_W_EMPTY + w_slot));
              mbar_arrive_cluster_relaxed(mapa_u32(bar(PB_W_EMPTY + w_slot), 1));
            }
            __syncwarp();
          } else {
            umma_commit2(bar(PB_W_EMPTY + w_slot), elected);
          }
        }   // previous chunk consumed (in both CTAs) once its MMAs complete
        if (++w_slot == ns) { w_slot = 0; ++w_use; }
        const long long tw0 = PT_NOW();
        (void)tw0;
        pwait(bar(PB_W_FULL + w_slot), w_use & 1u, 4u | ((uint32_t)tdbg << 8));   // (cluster-scope acquire would add CCTL.IVALL per wait)
        if (lane == 0) PT_ADD(9, PT_NOW() - tw0);
        tc_fence_after();
        return w_lo0 + (uint32_t)w_slot * (kPairSlotBytes >> 4);
      };
      auto chunk = [&](int kc, uint32_t d_tmem, uint32_t b_lo, uint32_t first_accumulates) {
#ifdef SVDLSTM_TC_TIMELINE
        const long long c0 = clock64();
#endif
        const uint32_t a = next_slot();
#ifdef SVDLSTM_TC_TIMELINE
        const long long c1 = clock64();
#endif
        if (kc == 64) umma2_f16_x<4>(d_tmem, a, a_hi64, b_lo, act_hi, idesc, first_accumulates, mma_el);
        else if (kc == 48) umma2_f16_x<3>(d_tmem, a, desc_hi(768u), b_lo, act_hi, idesc, first_accumulates, mma_el);
        else if (kc == 32) umma2_f16_x<2>(d_tmem, a, desc_hi(512u), b_lo, act_hi, idesc, first_accumulates, mma_el);
        else umma2_f16_x<1>(d_tmem, a, desc_hi(256u), b_lo, act_hi, idesc, first_accumulates, mma_el);
#ifdef SVDLSTM_TC_TIMELINE
        const long long c2 = clock64();
        if (lane == 0) { PT_ADD(10, c1 - c0); PT_ADD(11, c2 - c1); PT_ADD(12, 1); }
#endif
      };
      const int kin = p.kin, ru_pad = p.ru_pad;
      int in_s = 0;
      uint32_t in_ph = 0;
      uint32_t n_use = 0;   // uses of each S2 TMEM slot so far (both slots advance together)
#pragma unroll 1
      for (int t = 0; t < n_steps; ++t) {
        tdbg = t;
        if (dead) break;
        // ---- S1: [t_u(t) ; X(t-1)] = A1 . h(t-1)
        if (lane == 0) PT_STAMP(0);
        if (t > 0) {
          pwait(bar(PB_H_READY), (uint32_t)(t - 1) & 1u, 5u | ((uint32_t)tdbg << 8));
          fence_proxy_async();   // h(t-1) was written with st.async (generic proxy); the MMAs read it through the async proxy
          tc_fence_after();
        }
#pragma unroll 1
        for (int mt = 0; mt < p.n_mt; ++mt) {
          uint32_t b = h_lo0;
#pragma unroll 1
          for (int k0 = 0; k0 < H; k0 += 64) {
            chunk(64, tm_s1 + (uint32_t)mt * 128u, b, k0 > 0 ? 1u : 0u);
            b += kK64;
          }
        }
        if (lane == 0) PT_STAMP(1);
        umma_commit2(bar(PB_S1_FULL), elected);
        if (lane == 0) PT_STAMP(2);
        if (t == T) break;
        // ---- S2: z = A2in . in(t) (early, independent of h(t-1)) + A2u . t_u(t) (late)
        pwait(bar(PB_IN_FULL + in_s), in_ph, 6u | ((uint32_t)tdbg << 8));
        tc_fence_after();
        if (lane == 0) PT_STAMP(3);
        const uint32_t be0 = in_lo0 + (uint32_t)in_s * in_stage_lo;
#pragma unroll 1
        for (int ub = 0; ub < NUBC; ++ub)
#pragma unroll 1
          for (int gp = 0; gp < 2; ++gp) {
#pragma unroll 1
            for (int sl = 0; sl < 2; ++sl) {
              if (n_use >= 1u) {   // the epilogues of BOTH CTAs drained this slot's previous tile
                pwait(bar(PB_S2_EMPTY0 + sl), (n_use - 1u) & 1u, 7u | ((uint32_t)tdbg << 8));
                tc_fence_after();
              }
              uint32_t b = be0;
#pragma unroll 1
              for (int k0 = 0; k0 < kin; k0 += 64) {
                chunk(imin(64, kin - k0), tm_s2 + (uint32_t)sl * 128u, b, k0 > 0 ? 1u : 0u);
                b += kK64;
              }
            }
            if (ub == NUBC - 1 && gp == 1) {   // in(t) consumed
              umma_commit2(bar(PB_IN_EMPTY + in_s), elected);
              if (++in_s == nst) { in_s = 0; in_ph ^= 1u; }
            }
            if (lane == 0) PT_STAMP(4 + 2 * gp);   // early part of gate pair gp issued
            if (ub == 0 && gp == 0) {
              pwait(bar(PB_T_READY), (uint32_t)t & 1u, 8u | ((uint32_t)tdbg << 8));
              fence_proxy_async();
              tc_fence_after();
              if (lane == 0) PT_STAMP(8);
            }
#pragma unroll 1
            for (int sl = 0; sl < 2; ++sl) {
              uint32_t b = t_lo0;
#pragma unroll 1
              for (int k0 = 0; k0 < ru_pad; k0 += 64) {
                chunk(imin(64, ru_pad - k0), tm_s2 + (uint32_t)sl * 128u, b, 1u);
                b += kK64;
              }
              umma_commit2(bar(PB_S2_FULL0 + sl), elected);
            }
            if (lane == 0) PT_STAMP(5 + 2 * gp);   // late part of gate pair gp issued + committed
            ++n_use;
          }
      }
    } else {
      // ======================= peer CTA: forward "landed in my shared memory" to the leader's barriers ============
      if (lane == 0 && !(pp.off & 256)) {
        const uint32_t lead0 = mapa_u32(bar(PB_W_FULL), 0);
        int slot = 0;
        uint32_t use = 0;
        const long long total = (long long)T * n_fill_step + p.n_s1;
#pragma unroll 1
        for (long long i = 0; i < total; ++i) {
          tdbg = (int)(i / n_fill_step);
          if (dead) break;
          pwait(bar(PB_W_FULL + slot), use & 1u, 9u | ((uint32_t)tdbg << 8));
          if (dead) break;
          if (!(pp.off & 64)) mbar_arrive_cluster_relaxed(lead0 + 8u * (uint32_t)slot);
          if (++slot == ns) { slot = 0; ++use; }
        }
      }
      __syncwarp();
    }
  } else {
    // ======================= epilogue warps (512 threads: thread = TMEM lane (row) x 32 of the 128 columns) =========
    const int ew = warp - 4;
    const int q = ew & 3;                      // TMEM lane quarter (== warp % 4)
    const int cg = ew >> 2;                    // column group: sequences [32 cg, 32 cg + 32) of the pair tile
    const int row = q * 32 + lane;
    const uint32_t dst = (uint32_t)cg >> 1;    // CTA whose B-operand half holds these sequences
    const int lc0 = (cg & 1) * 32;             // first column inside that half
    const uint32_t tm_lane = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 32);   // + 128 per tile / slot
    float cst[NUBC][32];
#pragma unroll
    for (int u = 0; u < NUBC; ++u)
#pragma unroll
      for (int n = 0; n < 32; ++n) cst[u][n] = 0.f;
    float bia[NUBC][4];
#pragma unroll
    for (int u = 0; u < NUBC; ++u)
#pragma unroll
      for (int g = 0; g < 4; ++g) bia[u][g] = p.bias[(((size_t)rank * NUBC + u) * 4 + g) * 128 + row];
    // S1 plan of this thread, packed into one flag word (everything else is re-derived where it is used)
    int u_mt = 0, x_mt = 0;
    uint32_t fl = 0;   // 1: U rows loaded by this warp, 2: this row stored, 4: this row live, 8: X half present, 16: X rows loaded by this warp,
                       // 32: X row live, 64: X rows are the Dense top
    for (int mt = 0; mt < p.n_mt; ++mt) {
      const int k = p.half_kind[mt][rank], r0 = p.half_r0[mt][rank];
      if (k == PK_U) {
        u_mt = mt;
        if (r0 + q * 32 < p.ru_pad) fl |= 1u;
        if (r0 + row < p.ru_pad) fl |= 2u;
        if (r0 + row < p.ru) fl |= 4u;
      } else if (k == PK_X || k == PK_DENSE) {
        x_mt = mt;
        fl |= 8u;
        if (r0 + q * 32 < (k == PK_DENSE ? p.n_dense : p.x_pad)) fl |= 16u;
        if (r0 + row < (k == PK_DENSE ? p.n_dense : p.x_rows)) fl |= 32u;
        if (k == PK_DENSE) fl |= 64u;
      }
    }
    const uint32_t tb_u = mapa_u32(sbase + sp.tbuf, dst) + act_offset(p.half_r0[u_mt][rank] + row, lc0, 64);
    const uint32_t hb_addr = mapa_u32(sbase + sp.hbuf, dst) + act_offset((int)rank * (H / 2) + row, lc0, 64);   // + kRowBlk per own unit block
    const uint32_t lead_bar = mapa_u32(bar0, 0);   // the leader's barrier block
    const uint32_t dst_bar = mapa_u32(bar0, dst);  // the barrier block of the CTA this thread's operand rows go to
    uint32_t n_use = 0;   // S2 slot uses consumed so far
    for (int t = 0; t < n_steps; ++t) {
      tdbg = t;
      if (__any_sync(0xffffffffu, dead)) break;
      // ---- epilogue 1: S1 accumulators.  U rows -> f16 rows of the S2 B operand (both CTAs' halves) ----
      pwait(bar(PB_S1_FULL), (uint32_t)t & 1u, 11u | ((uint32_t)tdbg << 8));
      tc_fence_after();
      if (threadIdx.x == 128) PT_STAMP(16 + 16 * (int)rank + 0);
      // Exchange tracking (epilogue warp 0 of each CTA): arm the two tx-counted barriers of THIS CTA's operand halves for this step
      // (both earlier phases are complete: h(t-1) before this S1 pass, t_u(t-1) before the previous S2); on the peer, forward "my half
      // has landed" to the leader once it has.
      if (ew == 0 && lane == 0 && t < T) {
        mbar_expect_tx(bar(PB_T_READY), (uint32_t)p.ru_pad * 128u);
        mbar_expect_tx(bar(PB_H_READY), (uint32_t)H * 128u);
      }
      if ((fl & 1u) && t < T) {
        const uint32_t m = (fl & 4u) ? 0xFFFFFFFFu : 0u;   // padding rows (rank..rank_pad) must hold zeros
        // (TMEM reuse: the next S1 pass is issued after h(t) is complete, i.e. after every epilogue thread's E2 stores, which follow
        //  all of its TMEM loads of this step in program order)
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          uint32_t r[8];
          if (ep_ld) tmem_ld8(tm_lane + (uint32_t)u_mt * 128u + 8u * (uint32_t)j8, r);
          tmem_ld_wait();
          if ((fl & 2u) && ep_st)
            sta128_cluster(tb_u + 128u * (uint32_t)j8, dst_bar + 8u * PB_T_READY, pack_f16(__uint_as_float(r[0]), __uint_as_float(r[1])) & m,
                           pack_f16(__uint_as_float(r[2]), __uint_as_float(r[3])) & m, pack_f16(__uint_as_float(r[4]), __uint_as_float(r[5])) & m,
                           pack_f16(__uint_as_float(r[6]), __uint_as_float(r[7])) & m);
        }
      }
      if (ew == 0 && rank == 1 && t < T) {
        pwait(bar(PB_T_READY), (uint32_t)t & 1u, 17u | ((uint32_t)tdbg << 8));
        if (lane == 0) {
          fence_proxy_async();
          mbar_arrive_cluster_relaxed(lead_bar + 8u * PB_T_READY);
        }
        __syncwarp();
      }
      if (threadIdx.x == 128) PT_STAMP(16 + 16 * (int)rank + 1);
      // X rows: read after the t operand is published (off the critical path).  No TMEM hazard: the next S1 pass waits
      // for every warp's h(t), which each warp publishes after this read.
      if (t > 0 && (fl & 8u)) {
        if (fl & 64u) {   // Dense top: y(t-1)
          if (fl & 16u) {   // warp-uniform: the TMEM loads are .sync.aligned
            const int b_first = tile * 128 + cg * 32;
            const int o = p.half_r0[x_mt][rank] + row;
            const bool y_live = (fl & 32u) != 0;
            const int y_valid = y_live ? p.B - b_first : 0;
            float* yp = p.y + ((size_t)b_first * T + (t - 1)) * p.n_dense + (y_live ? o : 0);
            const float y_bias = y_live ? p.dense_bias[o] : 0.f;
            const size_t y_seq_stride = (size_t)T * p.n_dense;
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              uint32_t r[8];
              if (ep_ld) tmem_ld8(tm_lane + (uint32_t)x_mt * 128u + 8u * (uint32_t)j8, r);
              tmem_ld_wait();
#pragma unroll
              for (int n = 0; n < 8; ++n)
                if (8 * j8 + n < y_valid) yp[(size_t)(8 * j8 + n) * y_seq_stride] = __uint_as_float(r[n]) + y_bias;
            }
          }
        } else {          // next layer's t_w(t-1): stage this CTA's rows for both sequence halves
          if (t > 1) pwait(bar(PB_X_STORED), (uint32_t)(t - 2) & 1u, 12u | ((uint32_t)tdbg << 8));   // staging buffer of step t-2 shipped
          if (fl & 16u) {
            const uint32_t xs_addr = sbase + sp.xstage + dst * kRowBlk + act_offset(row, lc0, 64);
            const uint32_t m = (fl & 32u) ? 0xFFFFFFFFu : 0u;
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              uint32_t r[8];
              if (ep_ld) tmem_ld8(tm_lane + (uint32_t)x_mt * 128u + 8u * (uint32_t)j8, r);
              tmem_ld_wait();
              if (ep_st) sts128(xs_addr + 128u * (uint32_t)j8, pack_f16(__uint_as_float(r[0]), __uint_as_float(r[1])) & m,
                     pack_f16(__uint_as_float(r[2]), __uint_as_float(r[3])) & m, pack_f16(__uint_as_float(r[4]), __uint_as_float(r[5])) & m,
                     pack_f16(__uint_as_float(r[6]), __uint_as_float(r[7])) & m);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(PB_X_STAGED));
        }
      }
      if (t == T) break;

      // ---- epilogue 2: gates + cell update of the own cells, tile by tile through the two TMEM slots -----------------
      const uint32_t tz = tm_lane + 256u;
#pragma unroll
      for (int ub = 0; ub < NUBC; ++ub) {
        float ig[32];
        // gate i (slot 0)
        pwait(bar(PB_S2_FULL0), n_use & 1u, 13u | ((uint32_t)tdbg << 8));
        tc_fence_after();
        if (threadIdx.x == 128) PT_STAMP(16 + 16 * (int)rank + 2);
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          uint32_t z[8];
          if (ep_ld) tmem_ld8(tz + 8u * (uint32_t)j8, z);
          tmem_ld_wait();
#pragma unroll
          for (int n = 0; n < 8; ++n) ig[8 * j8 + n] = fmaf(0.5f, tanh_approx(__uint_as_float(z[n]) + bia[ub][0]), 0.5f);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(lead_bar + 8u * PB_S2_EMPTY0);
        // gate g (slot 1)
        pwait(bar(PB_S2_FULL1), n_use & 1u, 14u | ((uint32_t)tdbg << 8));
        tc_fence_after();
        if (threadIdx.x == 128) PT_STAMP(16 + 16 * (int)rank + 3);
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          uint32_t z[8];
          if (ep_ld) tmem_ld8(tz + 128u + 8u * (uint32_t)j8, z);
          tmem_ld_wait();
#pragma unroll
          for (int n = 0; n < 8; ++n) ig[8 * j8 + n] *= tanh_approx(__uint_as_float(z[n]) + bia[ub][1]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(lead_bar + 8u * PB_S2_EMPTY1);
        ++n_use;
        // gate f (slot 0): c = f c + i g
        pwait(bar(PB_S2_FULL0), n_use & 1u, 15u | ((uint32_t)tdbg << 8));
        tc_fence_after();
        if (threadIdx.x == 128) PT_STAMP(16 + 16 * (int)rank + 4);
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          uint32_t z[8];
          if (ep_ld) tmem_ld8(tz + 8u * (uint32_t)j8, z);
          tmem_ld_wait();
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            const float fg = fmaf(0.5f, tanh_approx(__uint_as_float(z[n]) + bia[ub][2]), 0.5f);
            cst[ub][8 * j8 + n] = fmaf(fg, cst[ub][8 * j8 + n], ig[8 * j8 + n]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(lead_bar + 8u * PB_S2_EMPTY0);
        // gate o (slot 1): h = o tanh(c) -> B-operand image of the CTA that owns these sequences
        pwait(bar(PB_S2_FULL1), n_use & 1u, 16u | ((uint32_t)tdbg << 8));
        tc_fence_after();
        if (threadIdx.x == 128) PT_STAMP(16 + 16 * (int)rank + 5);
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          uint32_t z[8];
          if (ep_ld) tmem_ld8(tz + 128u + 8u * (uint32_t)j8, z);
          tmem_ld_wait();
          if (j8 == 3) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(lead_bar + 8u * PB_S2_EMPTY1);
          }
          float hv[8];
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            const float og = fmaf(0.5f, tanh_approx(__uint_as_float(z[n]) + bia[ub][3]), 0.5f);
            hv[n] = og * tanh_approx(cst[ub][8 * j8 + n]);
          }
          if (ep_st) sta128_cluster(hb_addr + (uint32_t)ub * kRowBlk + 128u * (uint32_t)j8, dst_bar + 8u * PB_H_READY, pack_f16(hv[0], hv[1]),
                         pack_f16(hv[2], hv[3]), pack_f16(hv[4], hv[5]), pack_f16(hv[6], hv[7]));
        }
        ++n_use;
      }
      if (ew == 0 && rank == 1) {
        pwait(bar(PB_H_READY), (uint32_t)t & 1u, 18u | ((uint32_t)tdbg << 8));
        if (lane == 0) {
          fence_proxy_async();
          mbar_arrive_cluster_relaxed(lead_bar + 8u * PB_H_READY);
        }
        __syncwarp();
      }
      if (threadIdx.x == 128) PT_STAMP(16 + 16 * (int)rank + 6);
    }
  }

  // ---- teardown: the leader's MMAs read the peer's shared memory / write its TMEM until the very end ----------------
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// packing of the per-rank weight stream images
// ------------------------------------------------------------------------------------------------
__global__ void pack_pair_kernel(const PairChunk* __restrict__ chunks, Block bw, Block bu, Block bw_next, int H, int D,
                                 const float* __restrict__ dense_k, int n_dense, int n_out, __half* __restrict__ img) {
  const PairChunk c = chunks[blockIdx.x];
  if (c.kind == PK_NONE) return;
  __half* out = img + c.byte_off / 2;
  const int total = c.rows * c.kc;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int row = idx / c.kc, k = idx - row * c.kc;
    const int kk = c.k0 + k;
    float v = 0.f;
    if (c.kind == PK_U) {
      const int j = c.r0 + row;
      if (j < bu.rank) v = block_left(bu, kk, j);
    } else if (c.kind == PK_X) {
      const int j = c.r0 + row;
      if (j < bw_next.rank) v = block_left(bw_next, kk, j);
    } else if (c.kind == PK_DENSE) {
      const int j = c.r0 + row;
      if (j < n_dense) v = dense_k[(size_t)kk * n_out + j];
    } else {
      const int gate = (c.gate == 1) ? 2 : (c.gate == 2) ? 1 : c.gate;   // tile slots i, g(cell), f, o; Keras columns i, f, c, o
      const int n = gate * H + c.cell0 + row;
      if (c.kind == PK_S2_LATE) {
        if (kk < bu.rank) v = block_right(bu, kk, n);
      } else if (c.kind == PK_S2_EARLY) {
        if (kk < bw.rank) v = block_right(bw, kk, n);
      } else if (kk < D) {   // layer 0: the dense D x 4H product (L_w sigma_w) R_w
        float acc = 0.f;
        for (int j = 0; j < bw.rank; ++j) acc = fmaf(block_left(bw, kk, j), block_right(bw, j, n), acc);
        v = acc;
      }
      if (gate != 2) v *= 0.5f;
    }
    const size_t off = (size_t)(row / 8) * ((size_t)c.kc * 16) + (size_t)(k / 8) * 128 + (size_t)(row % 8) * 16 + (size_t)(k % 8) * 2;
    out[off / 2] = __float2half_rn(v);
  }
}
