// K1b: tcgen05 tensor-core persistent recurrent kernel (reduced-precision engine, FP16 operands, FP32
// accumulation and cell state).
//
// One CTA owns a tile of N=32 sequences for all T steps of ONE layer; layers run back to back and hand
// the hidden sequence over through HBM in the exact shared-memory operand image the next layer's MMAs
// consume (one 1-D bulk copy per step).  Per step the low-rank cell is two dependent thin contractions
// on the 5th-gen tensor cores, SWAP-AB so that the weight dimension fills the 128-row MMA and the batch
// tile is the (small) N:
//   S1u: t_u[r_u x N]  = (L_u sigma_u)^T [r_u x H]   . h(t-1) [H x N]      (+ Dense-top rows: y(t-1))
//   S1w: t_w[r_w x N]  = (L_w sigma_w)^T [r_w x Hin] . in(t)  [Hin x N]    (layers >= 1; issued one
//                                                                           step early: no recurrence)
//   S2 : z [4H x N]    = [R_u ; R_w]^T   [4H x K2]   . [t_u ; t_w | x(t)]  (layer 0: x enters S2 directly
//                                                                           through the dense 16x4H W);
//        issued in two passes: the input part (t_w | x) BEFORE t_u is ready, the recurrent part after
// Accumulators live in TMEM (S1: <= 5 x 32 columns; S2: two buffers of 4 gates x 32 columns, so the
// gate/cell epilogue of one 128-unit block overlaps the MMAs of the next).  The weight factors are one
// "weight stream": a fixed sequence of K-major 128-row chunks (<= 16 KB, <= 4 MMAs each) consumed in
// the same order every step.  If the stream fits shared memory it is loaded once and stays RESIDENT;
// otherwise it is re-streamed from L2 every step through a ring of 16 KB slots (full/empty mbarriers),
// which is what lets ranks up to 256 (1.3 MB of factors per layer) run on the tensor cores at all.
// Warp roles: warp 0 = weight streamer, warp 1 = MMA issuer (whole warp converged, one elected lane
// issues: the divergent single-thread form costs 2-3x more cycles per tcgen05.mma, see
// scripts/ubench_umma.cu), warp 2 = hidden-sequence stores, warp 3 = input ring (bulk copies), warps 4-11 =
// epilogue (TMEM -> registers -> activations -> FP16 operand in smem; two warps per TMEM lane quarter,
// 16 columns each).  Cell state c stays in FP32 registers for all T steps.  The Dense(1) top of the last
// layer rides in spare rows of the S1u tile (its output for step t-1 falls out of step t's first MMA chain).
//
// Replaces SingularLSTMCell.call / ReducedLSTMCell.call + backend.rnn for large batches
// (reference code/svd_classes_v3.py:116-145, 317-328, 405-434) at reduced precision; the FP32 engines
// remain the parity path.
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace svdlstm {

namespace {

constexpr int kInStagesMax = 3;      // max input prefetch ring depth
constexpr int kEpiWarps = 8;
constexpr int kTcThreads = 32 * (4 + kEpiWarps);
constexpr int kMaxWSlots = 12;
constexpr uint32_t kSlotBytes = 32768;   // preferred ring-slot size (two 16 KB chunks per hand-over); 16 KB when fewer than 4 would fit
constexpr int kMaxUB = 8;            // unit blocks of 128 (H <= 1024; above 512 with 16 epilogue warps, see tc_layer_body)
constexpr uint32_t kSmemCap = 227u * 1024u;
constexpr int kDbgPerLayer = 64 * 16 + 256;   // timeline: 16 stamps x 64 steps + per-chunk issue stamps of step 20

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped launch (CUDA error), never as a hung GPU.  The spin
// loop lives out of line: the kernel has ~40 wait sites and its hot code must stay inside the instruction cache.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FFu) == 0 && clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) { asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      ".reg .b32 rx;\n\t"
      "elect.sync rx|q, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, q;\n\t"
      "}\n"
      : "=r"(pred));
  return pred;
}
// tcgen05.mma (kind::f16: f16 x f16 -> fp32) is executed by the whole (converged) warp; only the elected lane issues.
// Descriptors are given as (lo, hi) words: stepping K is one add on lo.
// NM (1..4) back-to-back MMAs of one weight chunk in ONE asm block: the K step of both descriptors is an
// immediate add on the lo word (A +256 B; B +2 k-groups = NS*32 B per K=16) and the predicates are set up once per
// chunk.  NS = sequences per tile (MMA N): 32 or 64.
#define SVD_UMMA_HEAD                                                   \
  "{\n\t"                                                               \
  ".reg .pred p, q, one;\n\t"                                           \
  ".reg .b64 da, db;\n\t"                                               \
  ".reg .b32 ta, tb;\n\t"                                               \
  "setp.ne.b32 q, %7, 0;\n\t"                                           \
  "setp.ne.b32 p, %6, 0;\n\t"                                           \
  "setp.eq.b32 one, 0, 0;\n\t"                                          \
  "mov.b64 da, {%1, %2};\n\t"                                           \
  "mov.b64 db, {%3, %4};\n\t"                                           \
  "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
#define SVD_UMMA_NEXT(aoff, boff)                                       \
  "add.u32 ta, %1, " #aoff ";\n\t"                                      \
  "add.u32 tb, %3, " #boff ";\n\t"                                      \
  "mov.b64 da, {ta, %2};\n\t"                                           \
  "mov.b64 db, {tb, %4};\n\t"                                           \
  "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, one;\n\t"
#define SVD_UMMA_OPERANDS                                                                                              \
  ::"r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(first_accumulates), "r"(elected) : "memory"
template <int NM, int NS>
__device__ __forceinline__ void umma_f16_x(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                           uint32_t first_accumulates, uint32_t elected) {
  static_assert(NM >= 1 && NM <= 4 && (NS == 32 || NS == 64), "1..4 MMAs per chunk, 32 or 64 sequences");
  if constexpr (NM == 1 && NS == 32) asm volatile(SVD_UMMA_HEAD "}\n" SVD_UMMA_OPERANDS);
  if constexpr (NM == 2 && NS == 32) asm volatile(SVD_UMMA_HEAD SVD_UMMA_NEXT(16, 64) "}\n" SVD_UMMA_OPERANDS);
  if constexpr (NM == 3 && NS == 32) asm volatile(SVD_UMMA_HEAD SVD_UMMA_NEXT(16, 64) SVD_UMMA_NEXT(32, 128) "}\n" SVD_UMMA_OPERANDS);
  if constexpr (NM == 4 && NS == 32) asm volatile(SVD_UMMA_HEAD SVD_UMMA_NEXT(16, 64) SVD_UMMA_NEXT(32, 128) SVD_UMMA_NEXT(48, 192) "}\n" SVD_UMMA_OPERANDS);
  if constexpr (NM == 1 && NS == 64) asm volatile(SVD_UMMA_HEAD "}\n" SVD_UMMA_OPERANDS);
  if constexpr (NM == 2 && NS == 64) asm volatile(SVD_UMMA_HEAD SVD_UMMA_NEXT(16, 128) "}\n" SVD_UMMA_OPERANDS);
  if constexpr (NM == 3 && NS == 64) asm volatile(SVD_UMMA_HEAD SVD_UMMA_NEXT(16, 128) SVD_UMMA_NEXT(32, 256) "}\n" SVD_UMMA_OPERANDS);
  if constexpr (NM == 4 && NS == 64) asm volatile(SVD_UMMA_HEAD SVD_UMMA_NEXT(16, 128) SVD_UMMA_NEXT(32, 256) SVD_UMMA_NEXT(48, 384) "}\n" SVD_UMMA_OPERANDS);
}
// The 4 gate tiles of one unit block in ONE asm block (resident mode, one K chunk of NM MMAs per tile): gate g's
// accumulator is D + NS g columns, its weight chunk follows gate g-1's (+NM*4096 bytes), the B operand is shared.
// All steps are immediates, so the 4*NM MMAs cost a few uniform instructions each and no branches.
#define SVD_UMMA_G(goff_a, goff_d, koff_a, koff_b, PRED)                \
  "add.u32 ta, %1, " #goff_a "+" #koff_a ";\n\t"                        \
  "add.u32 tb, %3, " #koff_b ";\n\t"                                    \
  "add.u32 td, %0, " #goff_d ";\n\t"                                    \
  "mov.b64 da, {ta, %2};\n\t"                                           \
  "mov.b64 db, {tb, %4};\n\t"                                           \
  "@q tcgen05.mma.cta_group::1.kind::f16 [td], da, db, %5, " PRED ";\n\t"
#define SVD_UMMA_G_HEAD                                                 \
  "{\n\t"                                                               \
  ".reg .pred p, q, one;\n\t"                                           \
  ".reg .b64 da, db;\n\t"                                               \
  ".reg .b32 ta, tb, td;\n\t"                                           \
  "setp.ne.b32 q, %7, 0;\n\t"                                           \
  "setp.ne.b32 p, %6, 0;\n\t"                                           \
  "setp.eq.b32 one, 0, 0;\n\t"
template <int NM, int NS>
__device__ __forceinline__ void umma_f16_gates(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                               uint32_t first_accumulates, uint32_t elected) {
  static_assert(NM >= 1 && NM <= 4 && (NS == 32 || NS == 64), "1..4 MMAs per gate tile, 32 or 64 sequences");
  if constexpr (NM == 1 && NS == 32) asm volatile(SVD_UMMA_G_HEAD SVD_UMMA_G(0, 0, 0, 0, "p") SVD_UMMA_G(256, 32, 0, 0, "p") SVD_UMMA_G(512, 64, 0, 0, "p") SVD_UMMA_G(768, 96, 0, 0, "p") "}\n" SVD_UMMA_OPERANDS);
  if constexpr (NM == 2 && NS == 32) asm volatile(SVD_UMMA_G_HEAD SVD_UMMA_G(0, 0, 0, 0, "p") SVD_UMMA_G(0, 0, 16, 64, "one") SVD_UMMA_G(512, 32, 0, 0, "p") SVD_UMMA_G(512, 32, 16, 64, "one") SVD_UMMA_G(1024, 64, 0, 0, "p") SVD_UMMA_G(1024, 64, 16, 64, "one") SVD_UMMA_G(1536, 96, 0, 0, "p") SVD_UMMA_G(1536, 96, 16, 64, "one") "}\n" SVD_UMMA_OPERANDS);
  if constexpr (NM == 3 && NS == 32) asm volatile(SVD_UMMA_G_HEAD SVD_UMMA_G(0, 0, 0, 0, "p") SVD_UMMA_G(0, 0, 16, 64, "one") SVD_UMMA_G(0, 0, 32, 128, "one") SVD_UMMA_G(768, 32, 0, 0, "p") SVD_UMMA_G(768, 32, 16, 64, "one") SVD_UMMA_G(768, 32, 32, 128, "one") SVD_UMMA_G(1536, 64, 0, 0, "p") SVD_UMMA_G(1536, 64, 16, 64, "one") SVD_UMMA_G(1536, 64, 32, 128, "one") SVD_UMMA_G(2304, 96, 0, 0, "p") SVD_UMMA_G(2304, 96, 16, 64, "one") SVD_UMMA_G(2304, 96, 32, 128, "one") "}\n" SVD_UMMA_OPERANDS);
  if constexpr (NM == 4 && NS == 32) asm volatile(SVD_UMMA_G_HEAD SVD_UMMA_G(0, 0, 0, 0, "p") SVD_UMMA_G(0, 0, 16, 64, "one") SVD_UMMA_G(0, 0, 32, 128, "one") SVD_UMMA_G(0, 0, 48, 192, "one") SVD_UMMA_G(1024, 32, 0, 0, "p") SVD_UMMA_G(1024, 32, 16, 64, "one") SVD_UMMA_G(1024, 32, 32, 128, "one") SVD_UMMA_G(1024, 32, 48, 192, "one") SVD_UMMA_G(2048, 64, 0, 0, "p") SVD_UMMA_G(2048, 64, 16, 64, "one") SVD_UMMA_G(2048, 64, 32, 128, "one") SVD_UMMA_G(2048, 64, 48, 192, "one") SVD_UMMA_G(3072, 96, 0, 0, "p") SVD_UMMA_G(3072, 96, 16, 64, "one") SVD_UMMA_G(3072, 96, 32, 128, "one") SVD_UMMA_G(3072, 96, 48, 192, "one") "}\n" SVD_UMMA_OPERANDS);
  if constexpr (NM == 1 && NS == 64) asm volatile(SVD_UMMA_G_HEAD SVD_UMMA_G(0, 0, 0, 0, "p") SVD_UMMA_G(256, 64, 0, 0, "p") SVD_UMMA_G(512, 128, 0, 0, "p") SVD_UMMA_G(768, 192, 0, 0, "p") "}\n" SVD_UMMA_OPERANDS);
  if constexpr (NM == 2 && NS == 64) asm volatile(SVD_UMMA_G_HEAD SVD_UMMA_G(0, 0, 0, 0, "p") SVD_UMMA_G(0, 0, 16, 128, "one") SVD_UMMA_G(512, 64, 0, 0, "p") SVD_UMMA_G(512, 64, 16, 128, "one") SVD_UMMA_G(1024, 128, 0, 0, "p") SVD_UMMA_G(1024, 128, 16, 128, "one") SVD_UMMA_G(1536, 192, 0, 0, "p") SVD_UMMA_G(1536, 192, 16, 128, "one") "}\n" SVD_UMMA_OPERANDS);
  if constexpr (NM == 3 && NS == 64) asm volatile(SVD_UMMA_G_HEAD SVD_UMMA_G(0, 0, 0, 0, "p") SVD_UMMA_G(0, 0, 16, 128, "one") SVD_UMMA_G(0, 0, 32, 256, "one") SVD_UMMA_G(768, 64, 0, 0, "p") SVD_UMMA_G(768, 64, 16, 128, "one") SVD_UMMA_G(768, 64, 32, 256, "one") SVD_UMMA_G(1536, 128, 0, 0, "p") SVD_UMMA_G(1536, 128, 16, 128, "one") SVD_UMMA_G(1536, 128, 32, 256, "one") SVD_UMMA_G(2304, 192, 0, 0, "p") SVD_UMMA_G(2304, 192, 16, 128, "one") SVD_UMMA_G(2304, 192, 32, 256, "one") "}\n" SVD_UMMA_OPERANDS);
  if constexpr (NM == 4 && NS == 64) asm volatile(SVD_UMMA_G_HEAD SVD_UMMA_G(0, 0, 0, 0, "p") SVD_UMMA_G(0, 0, 16, 128, "one") SVD_UMMA_G(0, 0, 32, 256, "one") SVD_UMMA_G(0, 0, 48, 384, "one") SVD_UMMA_G(1024, 64, 0, 0, "p") SVD_UMMA_G(1024, 64, 16, 128, "one") SVD_UMMA_G(1024, 64, 32, 256, "one") SVD_UMMA_G(1024, 64, 48, 384, "one") SVD_UMMA_G(2048, 128, 0, 0, "p") SVD_UMMA_G(2048, 128, 16, 128, "one") SVD_UMMA_G(2048, 128, 32, 256, "one") SVD_UMMA_G(2048, 128, 48, 384, "one") SVD_UMMA_G(3072, 192, 0, 0, "p") SVD_UMMA_G(3072, 192, 16, 128, "one") SVD_UMMA_G(3072, 192, 32, 256, "one") SVD_UMMA_G(3072, 192, 48, 384, "one") "}\n" SVD_UMMA_OPERANDS);
}
__device__ __forceinline__ void umma_commit(uint32_t bar, uint32_t elected) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}\n" ::"r"(bar),
      "r"(elected)
      : "memory");
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }   // bit 46: sm_100 descriptor

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ------------------------------------------------------------------------------------------------
// operand images (SWIZZLE_NONE canonical layouts, cute/arch/mma_sm100_desc.hpp bit layout)
//   K-major  A chunk [rows x kc]: elem(row,k) at (row/8)*(kc*16) + (k/8)*128 + (row%8)*16 + (k%8)*2     LBO=128, SBO=kc*16
//   MN-major B tile  [K x N=32] : elem(k,n)   at (k/8)*512 + (n/8)*128 + (k%8)*16 + (n%8)*2             LBO=512, SBO=128
// ------------------------------------------------------------------------------------------------
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {   // N = sequences per tile
  return (1u << 4)                      // c_format  F32
         | (0u << 7)                    // a_format  F16
         | (0u << 10)                   // b_format  F16
         | (0u << 15)                   // A K-major
         | (1u << 16)                   // B MN-major
         | ((uint32_t)(N >> 3) << 17)   // n_dim
         | ((uint32_t)(M >> 4) << 24);  // m_dim
}
//   MN-major B tile [K x NS]: elem(k,n) at (k/8)*(NS*16) + (n/8)*128 + (k%8)*16 + (n%8)*2        LBO = NS*16, SBO = 128
constexpr uint32_t kActSBO = 128;
__host__ __device__ inline uint32_t act_tile_bytes(int K, int ns) { return (uint32_t)(K / 8) * (uint32_t)ns * 16u; }
__host__ __device__ inline uint32_t act_offset(int k, int n, int ns) {
  return (uint32_t)(k / 8) * (uint32_t)ns * 16u + (uint32_t)(n / 8) * 128u + (uint32_t)(k % 8) * 16u + (uint32_t)(n % 8) * 2u;
}
__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline int imin(int a, int b) { return a < b ? a : b; }

__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ------------------------------------------------------------------------------------------------
// per-layer launch parameters
// ------------------------------------------------------------------------------------------------
struct PackChunk {
  uint32_t byte_off;   // offset of the chunk inside the weight stream image
  int16_t seg;         // 0 = W (A1w), 1 = U (A1u), 2 = A2
  int16_t rows, kc;    // chunk extent (rows multiple of 8, kc multiple of 16)
  int16_t r0, k0;      // seg 0/1: first row / first K;  seg 2: k0 = first K of the part
  int16_t ub, g, part; // seg 2
};

struct TcLayerParams {
  const uint8_t* wimg;      // weight stream image: [segment W][segment U][segment 2]
  const PackChunk* chunks;  // the same stream as a chunk table (offset, extent) in consumption order: [W | U | 2]
  int n_chunks_w, n_chunks_u, n_chunks_2;
  const uint32_t* slots;    // streamed: (offset >> 8) | (bytes >> 8) << 16 per ring slot fill, consecutive chunks packed up to 16 KB
  int n_slots_w, n_slots_u, n_slots_2;
  const float* bias;        // [nub][4][128]; gates i,f,o pre-scaled by 0.5 (sigmoid(x) = 0.5 tanh(x/2) + 0.5)
  const uint8_t* in_seq;    // activation tile images [cta][t], K = Kin
  int in_ring, out_ring;    // layer-pipelined launch: the hand-off images are RINGS of this many steps per tile (0 = one slot per step).
                            // The layers run a few steps apart, so a short ring stays L2-resident: the hidden sequence never goes to HBM.
  const float* x_raw;       // layer 0: the caller's x (B, T, D) float32 -- read and converted by the input warp itself (no pack_x pass,
  int x_dim;                //          no FP16 image of x in HBM); nullptr: bulk copies from in_seq
  const int* x_ready;       // with x_raw: steps of x uploaded so far (the host is still copying x in time slices), or nullptr
  uint8_t* out_seq;         // activation tile images [cta][t], K = H            (store_h)
  float* y;                 // (B, T, n_dense) fused Dense-top output            (n_dense > 0)
  const float* dense_bias;
  int H, Kin, T, B;         // H = units padded to a multiple of 128 (padded cells: zero weights, h = c = 0)
  int Hu;                   // the layer's true units (state arrays and the packers index with it)
  int ns;                   // sequences per CTA tile (MMA N): 32 or 64
  int ru, rw;               // true ranks
  int ru_pad, rw_pad;       // multiples of 16: K extents of t_u / t_w inside the S2 contraction (rw_pad = 0 on layer 0)
  int kx;                   // layer 0: K extent of the x part of S2 (= Kin); else 0
  int rows_u, rows_w;       // rows of the S1 A operands, multiples of 8 (rows_u includes the Dense-top rows)
  int n_dense;              // Dense-top outputs fused into S1u (0 = none)
  int has_s1w;              // layers >= 1 that receive h of the previous layer: they compute their own t_w = A1w . in(t)
  int store_h;              // hand h(t) tiles to the next layer through HBM
  // t_w hand-off: the NEXT layer's t_w(t) = (L_w sigma_w)^T h(t) shares its B operand with this layer's S1u, so it rides in
  // extra rows [x_r0, x_r0 + rx) of the S1u tiles (falling out of step t+1's first MMA chain, like the Dense top) and is handed
  // over -- already in the layout the next layer's S2 consumes -- instead of h: the next layer has no S1w segment to stream
  // and no H-row input tile; the hand-off image is rx_pad/H of the size.
  int store_x;              // this layer ships t_w tiles of the next layer (then store_h = 0)
  int rx, rx_pad, x_r0;     // rows of the next layer's t_w, padded to 16, first S1u row
  int early_part;           // what the early (input) part of S2 contracts with: 1 = dense W0 (layer 0: x), 2 = R_w (t_w tiles / own t_w)
  int in_stages;            // depth of the input prefetch ring
  int streaming, w_slots;   // weight stream: 0 = resident; 1 = ring of w_slots slots of slot_bytes
  uint32_t slot_bytes;
  uint32_t segw_bytes, segu_bytes, seg2_bytes;
  // initial_state / return_state (svd_classes_v3.py:393,433-434) of THIS layer, (B, H) float32 each, or nullptr.  h enters and leaves
  // as float32 but lives as FP16 between steps, exactly like every h(t) of this engine (h_n is that FP16 value): a sequence run in
  // time chunks with the state carried over is bit-identical to the unchunked run.
  const float* h0;
  const float* c0;
  float* h_n;
  float* c_n;
  long long* dbg;           // optional timeline buffer (CTA 0): 16 clock64 stamps per step; nullptr = off
};

struct TcSmemPlan {
  uint32_t w, hbuf, tbuf, inbuf, xbuf, bars, tmem_slot, ctab, total;
};

__host__ __device__ inline TcSmemPlan tc_plan(const TcLayerParams& p) {
  TcSmemPlan s;
  uint32_t off = 0;
  s.w = off;
  off += p.streaming ? (uint32_t)p.w_slots * p.slot_bytes : ((p.segw_bytes + p.segu_bytes + p.seg2_bytes + 1023u) & ~1023u);
  s.hbuf = off; off += act_tile_bytes(p.H, p.ns);
  s.tbuf = off; off += act_tile_bytes(p.ru_pad + p.rw_pad, p.ns);
  s.inbuf = off; off += (uint32_t)p.in_stages * act_tile_bytes(p.Kin, p.ns);
  s.xbuf = off; off += p.store_x ? act_tile_bytes(p.rx_pad, p.ns) : 0u;
  s.bars = off; off += 512;
  s.tmem_slot = off; off += 16;
  // streaming: the chunk table (offset, bytes in 256-byte units) lives in smem -- with a 227 KB carve-out there is no L1 to cache it
  s.ctab = off; off += p.streaming ? (uint32_t)(((p.n_slots_w + p.n_slots_u + p.n_slots_2) * 4 + 15) & ~15) : 0u;
  s.total = off;
  return s;
}

// barrier slots (8 B each) inside the `bars` block
enum {
  BAR_W_FULL = 0,
  BAR_W_EMPTY = kMaxWSlots,
  BAR_IN_FULL = 2 * kMaxWSlots,
  BAR_IN_EMPTY = BAR_IN_FULL + kInStagesMax,
  BAR_S1_FULL = BAR_IN_EMPTY + kInStagesMax,
  BAR_T_READY,
  BAR_S1W_FULL,                       // layers >= 1: t_w(t) accumulators complete (committed one step early)
  BAR_TW_READY,                       // ... and converted into the B operand
  BAR_S2_FULL0,
  BAR_S2_FULL1,
  BAR_S2_EMPTY0,
  BAR_S2_EMPTY1,
  BAR_H_READY,                        // kMaxUB of them
  BAR_H_DONE = BAR_H_READY + kMaxUB,
  BAR_H_STORED,
  BAR_COUNT
};
static_assert(BAR_COUNT * 8 <= 512, "barrier block too small");

// ------------------------------------------------------------------------------------------------
// the weight stream: ONE definition of the chunk order, shared by the packer (host), the streamer
// warp and the MMA warp.  f(bytes, ...) is called once per chunk in consumption order.
// ------------------------------------------------------------------------------------------------
// segment W: A1w, rows_w x Kin, for mt, for 64-wide K chunk
template <class P, class F>
__host__ __device__ inline void for_seg_w(const P& p, F&& f) {
#pragma unroll 1
  for (int r0 = 0; r0 < p.rows_w; r0 += 128) {
    const int rows = imin(128, p.rows_w - r0);
#pragma unroll 1
    for (int k0 = 0; k0 < p.Kin; k0 += 64) f((uint32_t)rows * 128u, r0, rows, k0);
  }
}
// segment U: A1u, rows_u x H, for 128-unit K block (so the MMAs of block kb can start as soon as the epilogue
// published that block of h), for mt, for the two 64-wide K chunks of the block
template <class P, class F>
__host__ __device__ inline void for_seg_u(const P& p, F&& f) {
#pragma unroll 1
  for (int kb = 0; kb < p.H / 128; ++kb)
#pragma unroll 1
    for (int r0 = 0; r0 < p.rows_u; r0 += 128) {
      const int rows = imin(128, p.rows_u - r0);
#pragma unroll 1
      for (int c2 = 0; c2 < 2; ++c2) f((uint32_t)rows * 128u, kb, r0, rows, kb * 128 + c2 * 64);
    }
}
// segment 2: A2 in TWO passes over the (unit block, gate) tiles.  Early pass = the part of the gate contraction that
// does not depend on h(t-1): K = rw_pad rows of t_w (layers >= 1, part 2) or K = kx rows of x(t) (layer 0, part 1);
// for the first unit block it is issued while the epilogue is still converting t_u.  Late pass = the recurrent part,
// K = ru_pad rows of t_u (part 0), accumulating on top.  Resident stream: per unit block, early pass of its 4 gate
// tiles, then late pass (each pass is one asm block).  Streamed: tile by tile, early chunks then late chunks (the E2
// epilogue of block 0 can then start after half of the step's chunks, and one tile's early chunks cover E1).
// Chunks of <= 64 K each.
template <class P, class F>
__host__ __device__ inline void for_seg_2(const P& p, F&& f) {
  const int ke = p.has_s1w ? p.rw_pad : p.kx;
  const int part_e = p.early_part;
  if (p.streaming) {   // streamed: tile by tile (early chunks, then late chunks of the same tile)
#pragma unroll 1
    for (int ub = 0; ub < p.H / 128; ++ub)
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
#pragma unroll 1
        for (int k0 = 0; k0 < ke; k0 += 64) {
          const int kc = imin(64, ke - k0);
          f((uint32_t)kc * 256u, ub, g, part_e, k0, kc);
        }
#pragma unroll 1
        for (int k0 = 0; k0 < p.ru_pad; k0 += 64) {
          const int kc = imin(64, p.ru_pad - k0);
          f((uint32_t)kc * 256u, ub, g, 0, k0, kc);
        }
      }
    return;
  }
#pragma unroll 1
  for (int ub = 0; ub < p.H / 128; ++ub) {
#pragma unroll 1
    for (int g = 0; g < 4; ++g)
#pragma unroll 1
      for (int k0 = 0; k0 < ke; k0 += 64) {
        const int kc = imin(64, ke - k0);
        f((uint32_t)kc * 256u, ub, g, part_e, k0, kc);
      }
#pragma unroll 1
    for (int g = 0; g < 4; ++g)
#pragma unroll 1
      for (int k0 = 0; k0 < p.ru_pad; k0 += 64) {
        const int kc = imin(64, p.ru_pad - k0);
        f((uint32_t)kc * 256u, ub, g, 0, k0, kc);
      }
  }
}

// Input warp of layer 0: x(t) of this tile straight from the caller's (B, T, D) float32 array (no separate packing pass, no FP16
// image of x in HBM).  Every lane converts the rows of its sequences (n = lane, lane + 32) to FP16 and scatters them into the
// MN-major operand tile.  The loads of step t+1 are issued -- all at once: one memory latency per step -- before this warp waits
// for the next ring stage, and the ring runs 2-3 steps ahead of the MMAs.  Rows k >= D of the tile keep the zeros of the initial
// fill.  Compiled only into the RAWX instantiations of the kernel (see tc_layer_body).
template <int NS>
__device__ __forceinline__ void tc_raw_x_loader(const float* __restrict__ x, const int* __restrict__ x_ready, int D, int B, int T, int cta,
                                            uint32_t inbuf, uint32_t in_tile, int nst, uint32_t bar_full0, uint32_t bar_empty0) {
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kLBO = (uint32_t)NS * 16u;     // bytes between consecutive k-groups (8 K rows) of an activation tile
  constexpr int SPL = NS / 32;                      // sequences per lane
  const bool fast = (D % 4) == 0 && D <= 16;        // <= 4 float4 per sequence held in registers
  float4 v[SPL][4];
  auto st16 = [&](uint32_t addr, float f) {
    const __half hv = __float2half_rn(f);
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(*reinterpret_cast<const unsigned short*>(&hv)) : "memory");
  };
  auto load_step = [&](int t) {
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
      const int b = cta * NS + lane + 32 * q;
      const float4* src = reinterpret_cast<const float4*>(x + ((size_t)b * T + t) * D);
#pragma unroll
      for (int c = 0; c < 4; ++c) v[q][c] = (b < B && 4 * c < D) ? __ldcg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);   // (L2 only: x may
    }                                                                                                               //  still be arriving by DMA)
  };
  // x uploaded in time slices while the kernel runs (svdlstm_forward_streamed_input): the host publishes, in stream order behind
  // each slice, how many steps have landed; re-read only when the next step is not covered by the last value seen
  int avail = x_ready != nullptr ? 0 : T;
  auto wait_for = [&](int t) {
    if (t < avail) return;
    if (lane == 0) {
      int a = ld_acquire_gpu(x_ready);
      while (a <= t) {
        __nanosleep(200);
        a = ld_acquire_gpu(x_ready);
      }
      avail = a;
    }
    avail = __shfl_sync(0xffffffffu, avail, 0);
  };
  int ld_s = 0;
  uint32_t ld_n = 0;
  wait_for(0);
  if (fast) load_step(0);
#pragma unroll 1
  for (int ld_t = 0; ld_t < T; ++ld_t) {
    if (ld_n > 0) {
      if (lane == 0) mbar_wait(bar_empty0 + 8u * (uint32_t)ld_s, (ld_n - 1u) & 1u);
      __syncwarp();
    }
    const uint32_t tile = inbuf + (uint32_t)ld_s * in_tile;
    if (fast) {
#pragma unroll
      for (int q = 0; q < SPL; ++q) {
        const int n = lane + 32 * q;
        const uint32_t col = tile + (uint32_t)(n / 8) * 128u + (uint32_t)(n % 8) * 2u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (4 * c < D) {
            const uint32_t a0 = col + (uint32_t)((4 * c) / 8) * kLBO + (uint32_t)((4 * c) % 8) * 16u;
            st16(a0, v[q][c].x);
            st16(a0 + 16u, v[q][c].y);
            st16(a0 + 32u, v[q][c].z);
            st16(a0 + 48u, v[q][c].w);
          }
        }
      }
    } else {
#pragma unroll 1
      for (int n = lane; n < NS; n += 32) {
        const int b = cta * NS + n;
        const float* src = x + ((size_t)b * T + ld_t) * D;
        const uint32_t col = tile + (uint32_t)(n / 8) * 128u + (uint32_t)(n % 8) * 2u;
#pragma unroll 4
        for (int kk = 0; kk < D; ++kk) st16(col + (uint32_t)(kk / 8) * kLBO + (uint32_t)(kk % 8) * 16u, b < B ? __ldcg(src + kk) : 0.f);
      }
    }
    fence_proxy_async();     // generic-proxy stores -> visible to the MMAs' async proxy
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_full0 + 8u * (uint32_t)ld_s);
    if (ld_t + 1 < T) wait_for(ld_t + 1);
    if (fast && ld_t + 1 < T) load_step(ld_t + 1);     // in flight while this warp waits for the next ring stage
    if (++ld_s == nst) { ld_s = 0; ++ld_n; }
  }
}

#ifdef SVDLSTM_TC_TIMELINE
#define TC_STAMP(slot) do { if (p.dbg != nullptr && stamp_cta && t < 64) p.dbg[t * 16 + (slot)] = clock64(); } while (0)
#define TC_CHUNK_STAMP() do { if (p.dbg != nullptr && stamp_cta && dbg_t == 20 && dbg_n < 250) p.dbg[64 * 16 + dbg_n++] = clock64(); } while (0)
#else
#define TC_STAMP(slot) do { } while (0)
#define TC_CHUNK_STAMP() do { } while (0)
#endif

// One CTA = one layer x one tile of NS sequences x all T steps.  prog_in / prog_out (layer-pipelined launch only): per-tile
// step counters in global memory through which the previous layer's CTA publishes -- and this CTA announces -- how many
// hidden-sequence tiles have landed in the hand-off image.
// EW = epilogue warps (8, or 16 for H > 512: the cell state of NUB x 128 cells x NS sequences lives in their registers).
// STATE = initial_state / return_state plumbing compiled in (a separate instantiation: even untaken, its extra live pointers cost
// the epilogue-bound ranks 20 %).
// RAWX = layer 0 may read the caller's float32 x itself (tc_raw_x_loader inlined into the input warp).  A separate instantiation:
// the epilogue warps sit at the register cap and ANY code compiled next to them perturbs their allocation -- with the loader
// compiled in, ranks >= 128 gain 3 % (no pack_x pass) while the epilogue-bound ranks <= 64 lose 6-18 % (measured, same box).
template <int NUB, bool STREAM, int NS, int EW = kEpiWarps, bool STATE = false, bool RAWX = false>
__device__ __forceinline__ void tc_layer_body(const TcLayerParams& p, const int cta, const int* prog_in, int* prog_out, const bool stamp_cta,
                                              int* cons_pub = nullptr, const int* cons_wait = nullptr) {
  constexpr int CPT = NS * 4 / EW;            // accumulator columns per epilogue thread (EW / 4 warps per TMEM lane quarter)
  constexpr int kTcThreads = 32 * (4 + EW);   // (shadow the 8-warp defaults of the file scope)
  constexpr int kEpiThreads = 32 * EW;
  static_assert(CPT == 8 || CPT == 16 || CPT == 32, "8, 16 or 32 accumulator columns per epilogue thread");
  constexpr uint32_t kRowBlk = (uint32_t)NS * 256u;   // bytes of 128 K-rows of an activation tile (16 k-groups)
  constexpr uint32_t kK64 = (uint32_t)NS * 8u;        // descriptor-lo step of 64 K-rows (8 k-groups of NS*16 bytes)
  constexpr int kS2Bufs = NS == 32 ? 2 : 1;   // S2 accumulator buffers that fit TMEM (4 gates x NS columns each)
  extern __shared__ __align__(1024) uint8_t smem[];
  const TcSmemPlan sp = tc_plan(p);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, T = p.T;
  constexpr int nub = NUB;
  const int nst = p.in_stages;
  const int ns = p.w_slots;
  const uint32_t in_tile = act_tile_bytes(p.Kin, NS), h_tile = act_tile_bytes(H, NS);
  const uint32_t bar0 = sbase + sp.bars;
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int n_steps = T + ((p.n_dense > 0 || p.store_x) ? 1 : 0);   // the Dense-top / t_w rows of step T-1 need one more S1u pass

  // ---- one-time setup ---------------------------------------------------------------------------
  // zero the activation buffers (h(-1) = 0; padded rows must be finite)
  for (uint32_t i = threadIdx.x * 16u; i < sp.bars - sp.hbuf; i += kTcThreads * 16u)
    *reinterpret_cast<uint4*>(smem + sp.hbuf + i) = make_uint4(0, 0, 0, 0);
  if (p.streaming)   // stale bytes behind short chunks feed unused accumulator rows only, but keep them finite
    for (uint32_t i = threadIdx.x * 16u; i < (uint32_t)ns * p.slot_bytes; i += kTcThreads * 16u)
      *reinterpret_cast<uint4*>(smem + sp.w + i) = make_uint4(0, 0, 0, 0);
  if (STREAM) {
    const int nc = p.n_slots_w + p.n_slots_u + p.n_slots_2;
    uint32_t* ctab = reinterpret_cast<uint32_t*>(smem + sp.ctab);
    for (int i = threadIdx.x; i < nc; i += kTcThreads) ctab[i] = p.slots[i];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxWSlots; ++s) {
      mbar_init(bar(BAR_W_FULL + s), 1);
      mbar_init(bar(BAR_W_EMPTY + s), 1);
    }
    for (int s = 0; s < kInStagesMax; ++s) {
      mbar_init(bar(BAR_IN_FULL + s), 1);
      mbar_init(bar(BAR_IN_EMPTY + s), 1);
    }
    mbar_init(bar(BAR_S1_FULL), 1);
    mbar_init(bar(BAR_T_READY), kEpiThreads);
    mbar_init(bar(BAR_S1W_FULL), 1);
    mbar_init(bar(BAR_TW_READY), kEpiThreads);
    mbar_init(bar(BAR_S2_FULL0), 1);
    mbar_init(bar(BAR_S2_FULL1), 1);
    mbar_init(bar(BAR_S2_EMPTY0), kEpiThreads);
    mbar_init(bar(BAR_S2_EMPTY1), kEpiThreads);
    for (int u = 0; u < kMaxUB; ++u) mbar_init(bar(BAR_H_READY + u), kEpiThreads);
    mbar_init(bar(BAR_H_DONE), kEpiThreads);
    mbar_init(bar(BAR_H_STORED), 1);
    fence_barrier_init();
  }
  if (STATE && p.h0 != nullptr) {   // h(-1) = initial state, as the FP16 B operand of the first S1u pass
    __syncthreads();       // (after the zero fill above)
    if (warp >= 4) {
      const int ew0 = warp - 4, row0 = (ew0 & 3) * 32 + lane, cc0 = (ew0 >> 2) * CPT;
      for (int ub = 0; ub < NUB; ++ub)
        for (int j8 = 0; j8 < CPT / 8; ++j8) {
          float v[8];
          for (int m = 0; m < 8; ++m) {
            const int b = cta * NS + cc0 + 8 * j8 + m;
            v[m] = (b < p.B && ub * 128 + row0 < p.Hu) ? p.h0[(size_t)b * p.Hu + ub * 128 + row0] : 0.f;
          }
          sts128(sbase + sp.hbuf + act_offset(ub * 128 + row0, cc0 + 8 * j8, NS), pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]),
                 pack_f16(v[6], v[7]));
        }
    }
  }
  if (warp == 1) tmem_alloc(sbase + sp.tmem_slot, 512);
  fence_proxy_async();   // generic-proxy zero fill -> visible to the async proxy (MMA / bulk copies)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + sp.tmem_slot);
  // TMEM columns (NS=32): [0,96) S1u (<= 3 row tiles), [96,160) S1w (<= 2), [192,320) S2 buffer 0 (4 gates x 32), [320,448) buffer 1
  //              (NS=64): [0,128) S1u (<= 2 row tiles), [128,256) S1w (<= 2), [256,512) the one S2 buffer (4 gates x 64)
  const uint32_t tm_s1u = tmem, tm_s1w = tmem + (NS == 32 ? 96u : 128u), tm_s2 = tmem + (NS == 32 ? 192u : 256u);

  if (warp == 0) {
    // ======================= weight streamer ====================================================
    if (lane == 0) {
      if (!STREAM) {
        const uint32_t wbytes = p.segw_bytes + p.segu_bytes + p.seg2_bytes;
        mbar_expect_tx(bar(BAR_W_FULL), wbytes);
#pragma unroll 1
        for (uint32_t o = 0; o < wbytes; o += 65536u) {
          const uint32_t n = wbytes - o < 65536u ? wbytes - o : 65536u;
          bulk_g2s(sbase + sp.w + o, p.wimg + o, n, bar(BAR_W_FULL));
        }
      } else {
        // walk the chunk table in consumption order: [W] then per step [U][2][W of the next step], finally [U] (Dense flush)
        int slot = 0;
        uint32_t use = 0;   // how many times the ring wrapped
        const int nw = p.n_slots_w, nu = p.n_slots_u, n2 = p.n_slots_2;
        int t = 0, ph = p.has_s1w ? 0 : 1;   // ph 0: W of step 0; then per step 1: U, 2: segment 2, 3: W of step t+1
#pragma unroll 1
        while (true) {
          int i0, n;
          if (ph == 0) { i0 = 0; n = nw; }
          else if (ph == 1) { i0 = nw; n = nu; }
          else if (ph == 2) { i0 = nw + nu; n = n2; }
          else { i0 = 0; n = (p.has_s1w && t + 1 < T) ? nw : 0; }
#pragma unroll 1
          for (int i = i0; i < i0 + n; ++i) {
            const uint32_t e = reinterpret_cast<const volatile uint32_t*>(smem + sp.ctab)[i];
            const uint32_t bytes = (e >> 16) << 8;
            if (use > 0) mbar_wait(bar(BAR_W_EMPTY + slot), (use - 1u) & 1u);
            mbar_expect_tx(bar(BAR_W_FULL + slot), bytes);
            bulk_g2s(sbase + sp.w + (uint32_t)slot * p.slot_bytes, p.wimg + ((size_t)(e & 0xFFFFu) << 8), bytes, bar(BAR_W_FULL + slot));
            if (++slot == ns) { slot = 0; ++use; }
          }
          if (ph == 0) ph = 1;
          else if (ph == 1) { if (t == T) break; ph = 2; }
          else if (ph == 2) ph = 3;
          else { if (++t >= n_steps) break; ph = 1; }
        }
      }
    }
  } else if (warp == 3) {
    // ======================= input ring ==========================================================================
    if (RAWX && p.x_raw != nullptr) {
      if constexpr (RAWX) tc_raw_x_loader<NS>(p.x_raw, p.x_ready, p.x_dim, p.B, T, cta, sbase + sp.inbuf, in_tile, nst, bar(BAR_IN_FULL), bar(BAR_IN_EMPTY));
    } else if (lane == 0) {
      // one bulk copy per step, as early as the ring allows
      const int iring = p.in_ring;
      const uint8_t* src = p.in_seq + (size_t)cta * (iring ? iring : T) * in_tile;
      int ld_s = 0;          // ring stage of the next tile to load
      uint32_t ld_n = 0;     // how many times that stage has been used
#pragma unroll 1
      for (int ld_t = 0; ld_t < T; ++ld_t) {
        if (prog_in != nullptr) {   // layer-pipelined launch: the previous layer must have published tile ld_t
          if (ld_acquire_gpu(prog_in) <= ld_t) {
            const long long t0 = clock64();
            while (ld_acquire_gpu(prog_in) <= ld_t)
              if (clock64() - t0 > 8000000000LL) __trap();
          }
          fence_proxy_async_all();   // the tile was written through the async proxy of another SM
        }
        if (ld_n > 0) {
          mbar_wait(bar(BAR_IN_EMPTY + ld_s), (ld_n - 1u) & 1u);
          // the MMAs that read the tile of step ld_t - nst have completed, so its copy out of the hand-off ring has too:
          // tell the producing layer that those ring slots may be overwritten
          if (cons_pub != nullptr) st_release_gpu(cons_pub, ld_t - nst + 1);
        }
        mbar_expect_tx(bar(BAR_IN_FULL + ld_s), in_tile);
        bulk_g2s(sbase + sp.inbuf + ld_s * in_tile, src + (size_t)(iring ? ld_t % iring : ld_t) * in_tile, in_tile, bar(BAR_IN_FULL + ld_s));
        if (++ld_s == nst) { ld_s = 0; ++ld_n; }
      }
    }
  } else if (warp == 2) {
    // ======================= hidden-sequence stores (hand-off to the next layer) ====================================
    if (lane == 0 && (p.store_h || p.store_x)) {
      const uint32_t o_tile = p.store_x ? act_tile_bytes(p.rx_pad, NS) : h_tile;   // what is handed over: t_w(t) of the next layer, or h(t)
      const uint32_t o_src = sbase + (p.store_x ? sp.xbuf : sp.hbuf);
      const int oring = p.out_ring;
      uint8_t* out = p.out_seq + (size_t)cta * (oring ? oring : T) * o_tile;
#pragma unroll 1
      for (int t = 0; t < T; ++t) {
        // tile t complete in smem -> ship it to the hand-off image, then let the epilogue overwrite the buffer
        mbar_wait(bar(BAR_H_DONE), (uint32_t)(t & 1));
        if (oring && t >= oring && cons_wait != nullptr && ld_acquire_gpu(cons_wait) < t - oring + 1) {   // ring slot still unread downstream
          const long long t0 = clock64();
          while (ld_acquire_gpu(cons_wait) < t - oring + 1)
            if (clock64() - t0 > 8000000000LL) __trap();
        }
        bulk_s2g(out + (size_t)(oring ? t % oring : t) * o_tile, o_src, o_tile);
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(bar(BAR_H_STORED));
        if (prog_out != nullptr) {   // publish: the tile is complete in global memory, then the step counter
          bulk_wait_all0();
          fence_proxy_async_all();
          st_release_gpu(prog_out, t + 1);
        }
      }
      bulk_wait_all0();
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (whole warp, converged; one elected lane issues) =============
    // This warp runs alone on its scheduler slot, so every instruction's latency is exposed: the issue
    // loops below are written to cost a handful of uniform-datapath instructions per tcgen05.mma
    // (descriptor words advanced by immediates inside one asm block per chunk, no per-MMA branches).
    const uint32_t elected = elect_one();
    const uint32_t idesc = make_idesc(128, NS);
    const uint32_t act_hi = desc_hi(kActSBO);
    constexpr uint32_t kActLBO = (uint32_t)NS * 16u;
    const uint32_t h_lo0 = desc_lo(sbase + sp.hbuf, kActLBO), t_lo0 = desc_lo(sbase + sp.tbuf, kActLBO);
    const uint32_t in_lo0 = desc_lo(sbase + sp.inbuf, kActLBO), in_stage_lo = in_tile >> 4;
    const uint32_t w_lo0 = desc_lo(sbase + sp.w, 128u);
    constexpr bool streaming = STREAM;
    const uint32_t a_hi64 = desc_hi(1024u);   // kc = 64 chunks
    const uint32_t slot_bytes = p.slot_bytes, slot_lo = p.slot_bytes >> 4;   // in registers: this warp pays every load's latency
    int w_slot = -1;            // streamed: ring slot the cursor is in (-1: none yet)
    uint32_t w_use = 0;         // how many times the ring wrapped
    uint32_t slot_left = 0;     // bytes of the current slot not yet consumed (consecutive chunks share a slot up to 16 KB)
    int dbg_n = 0, dbg_t = -1;
    (void)dbg_n;
    (void)dbg_t;
    uint32_t a_lo = w_lo0;   // descriptor word of the next chunk (resident: walks the image; streamed: walks the ring slot)
    if (!streaming) mbar_wait(bar(BAR_W_FULL), 0);
    // Where the next weight chunk lives.  Streamed: consecutive chunks of a segment are packed into 16 KB slots by the same
    // greedy rule the host used for the slot table; moving on to a new slot hands the previous one back to the streamer.
    auto next_chunk = [&](uint32_t bytes) -> uint32_t {
      if (streaming && slot_left < bytes) {
        if (w_slot >= 0) umma_commit(bar(BAR_W_EMPTY + w_slot), elected);
        if (++w_slot == ns) { w_slot = 0; ++w_use; }
        mbar_wait(bar(BAR_W_FULL + w_slot), w_use & 1u);
        tc_fence_after();
        slot_left = slot_bytes;
        a_lo = w_lo0 + (uint32_t)w_slot * slot_lo;
      }
      const uint32_t a = a_lo;
      a_lo += bytes >> 4;
      slot_left -= bytes;   // (resident: unused)
      return a;
    };
    auto new_segment = [&]() { slot_left = 0; };   // streamed: a segment starts in a fresh slot
    // one weight chunk = NM <= 4 MMAs (K = 16 each): A = chunk (K-major, SBO = 256*NM), B = activation rows
    // remainder chunk (K extent not a multiple of 64): nm < 4 single MMAs
    auto chunk = [&](int nm, uint32_t bytes, uint32_t a_hi, uint32_t d_tmem, uint32_t b_lo, uint32_t first_accumulates) {
      TC_CHUNK_STAMP();
      const uint32_t a = next_chunk(bytes);
      if (nm == 2) umma_f16_x<2, NS>(d_tmem, a, a_hi, b_lo, act_hi, idesc, first_accumulates, elected);        // ranks 17..32
      else if (nm == 1) umma_f16_x<1, NS>(d_tmem, a, a_hi, b_lo, act_hi, idesc, first_accumulates, elected);   // ranks <= 16, x(t)
      else umma_f16_x<3, NS>(d_tmem, a, a_hi, b_lo, act_hi, idesc, first_accumulates, elected);
    };
    auto chunk64 = [&](uint32_t bytes, uint32_t d_tmem, uint32_t b_lo, uint32_t first_accumulates) {   // the common case, no switch
      TC_CHUNK_STAMP();
      const uint32_t a = next_chunk(bytes);
      umma_f16_x<4, NS>(d_tmem, a, a_hi64, b_lo, act_hi, idesc, first_accumulates, elected);
    };
    // resident mode, one K chunk of nm MMAs per gate tile: the whole unit block (4 gates) goes out as one asm block
    auto gates_block = [&](int nm, uint32_t d_tmem, uint32_t b_lo, uint32_t first_accumulates) {
      const uint32_t a_hi = desc_hi((uint32_t)nm * 256u);
      switch (nm) {
        case 4: umma_f16_gates<4, NS>(d_tmem, a_lo, a_hi, b_lo, act_hi, idesc, first_accumulates, elected); break;
        case 3: umma_f16_gates<3, NS>(d_tmem, a_lo, a_hi, b_lo, act_hi, idesc, first_accumulates, elected); break;
        case 2: umma_f16_gates<2, NS>(d_tmem, a_lo, a_hi, b_lo, act_hi, idesc, first_accumulates, elected); break;
        default: umma_f16_gates<1, NS>(d_tmem, a_lo, a_hi, b_lo, act_hi, idesc, first_accumulates, elected); break;
      }
      a_lo += (uint32_t)nm * 1024u;   // 4 chunks of nm * 4096 bytes
    };
    const int n_in_chunks = p.Kin >> 6;
    const int ke = p.has_s1w ? p.rw_pad : p.kx;                 // early (input) part of S2
    const int ke_full = ke >> 6, ke_rem = (ke & 63) >> 4;
    const int ku_full = p.ru_pad >> 6, ku_rem = (p.ru_pad & 63) >> 4;   // late (recurrent) part
    const uint32_t tw_lo0 = t_lo0 + (uint32_t)(p.ru_pad >> 3) * (uint32_t)NS;    // t_w rows follow the t_u rows in the t buffer
    const int ke_one = ke <= 64 ? ke >> 4 : 0, ku_one = p.ru_pad <= 64 ? p.ru_pad >> 4 : 0;   // MMAs per tile if it is ONE chunk
    int in_s = 0;            // input ring stage of step t (layer 0) / of the next S1w (layers >= 1)
    uint32_t in_ph = 0;      // its phase bit
    auto issue_s1w = [&]() {   // t_w(tt) = A1w . in(tt), tt = the step the ring cursor points at
      mbar_wait(bar(BAR_IN_FULL + in_s), in_ph);
      tc_fence_after();
      const uint32_t in_lo = in_lo0 + (uint32_t)in_s * in_stage_lo;
      if (!streaming) a_lo = w_lo0;
      new_segment();
      uint32_t d = tm_s1w;
#pragma unroll 1
      for (int r0 = 0; r0 < p.rows_w; r0 += 128) {
        const uint32_t bytes = (uint32_t)imin(128, p.rows_w - r0) * 128u;
        uint32_t b = in_lo;
#pragma unroll 1
        for (int c = 0; c < n_in_chunks; ++c) {
          chunk64(bytes, d, b, c > 0 ? 1u : 0u);
          b += kK64;   // 64 K rows of the activation tile
        }
        d += (uint32_t)NS;
      }
      umma_commit(bar(BAR_IN_EMPTY + in_s), elected);   // in(tt) is consumed once these MMAs complete
      umma_commit(bar(BAR_S1W_FULL), elected);
      if (++in_s == nst) { in_s = 0; in_ph ^= 1u; }
    };
    uint32_t s2_use = 0;   // global count of S2 accumulator-buffer uses (buffers alternate)
    if (p.has_s1w) issue_s1w();
#pragma unroll 1
    for (int t = 0; t < n_steps; ++t) {
      dbg_t = t;
      // ---- S1u: [t_u ; y] = A1u . h(t-1), K block by K block as the epilogue publishes h(t-1)
      if (!streaming) a_lo = w_lo0 + (p.segw_bytes >> 4);
      new_segment();
      {
        uint32_t b = h_lo0;
#pragma unroll 1
        for (int kb = 0; kb < nub; ++kb) {
          if (t > 0) {
            mbar_wait(bar(BAR_H_READY + kb), (uint32_t)(t - 1) & 1u);
            tc_fence_after();
          }
          uint32_t d = tm_s1u;
#pragma unroll 1
          for (int r0 = 0; r0 < p.rows_u; r0 += 128) {
            const uint32_t bytes = (uint32_t)imin(128, p.rows_u - r0) * 128u;
            chunk64(bytes, d, b, kb > 0 ? 1u : 0u);
            chunk64(bytes, d, b + kK64, 1u);
            d += (uint32_t)NS;
          }
          b += 2u * kK64;   // next 128 K rows of h
        }
      }
      umma_commit(bar(BAR_S1_FULL), elected);
      TC_STAMP(2);   // MMA: S1 issued + committed
      if (t == T) break;
      // ---- S2 early pass: z = A2in . (t_w(t) | x(t)) -- nothing here depends on h(t-1)
      uint32_t be0;
      if (p.has_s1w) {
        mbar_wait(bar(BAR_TW_READY), (uint32_t)t & 1u);
        be0 = tw_lo0;
      } else {
        mbar_wait(bar(BAR_IN_FULL + in_s), in_ph);
        be0 = in_lo0 + (uint32_t)in_s * in_stage_lo;
      }
      tc_fence_after();
      new_segment();
#pragma unroll 1
      for (int ub = 0; ub < nub; ++ub) {
        // NS=32: two whole-block buffers (4 gate tiles each), used alternately.  NS=64: ONE block of TMEM, split into
        // half-buffers A = tiles (i, g) and B = tiles (f, o) that the epilogue drains -- and hands back -- separately.
        const uint32_t use = s2_use + (uint32_t)ub;
        const uint32_t buf = kS2Bufs == 2 ? (use & 1u) : 0u, turn = kS2Bufs == 2 ? (use >> 1) : use;
        if (turn >= 1u) {   // wait until the epilogue drained this TMEM buffer (its previous use)
          mbar_wait(bar(BAR_S2_EMPTY0 + buf), (turn - 1u) & 1u);
          if (kS2Bufs == 1 && !streaming) mbar_wait(bar(BAR_S2_EMPTY1), (turn - 1u) & 1u);   // resident passes touch all 4 tiles
          tc_fence_after();
        }
        const uint32_t d0 = tm_s2 + buf * (uint32_t)(4 * NS);
        if (streaming) {
          // tile by tile: early chunks (input part), then late chunks (recurrent part; the first one waits for t_u)
          uint32_t d = d0;
#pragma unroll 1
          for (int g = 0; g < 4; ++g) {
            if (kS2Bufs == 1 && g == 2 && turn >= 1u) {   // half-buffer B (tiles f, o) of the previous block drained?
              mbar_wait(bar(BAR_S2_EMPTY1), (turn - 1u) & 1u);
              tc_fence_after();
            }
            uint32_t b = be0;
            uint32_t acc = 0u;
#pragma unroll 1
            for (int c = 0; c < ke_full; ++c) {
              chunk64(16384u, d, b, acc);
              acc = 1u;
              b += kK64;
            }
            if (ke_rem) chunk(ke_rem, (uint32_t)ke_rem * 4096u, desc_hi((uint32_t)ke_rem * 256u), d, b, acc);
            if (ub == nub - 1 && g == 3 && !p.has_s1w) {   // layer 0: x(t) fully consumed
              umma_commit(bar(BAR_IN_EMPTY + in_s), elected);
              if (++in_s == nst) { in_s = 0; in_ph ^= 1u; }
            }
            if (ub == 0 && g == 0) {
              mbar_wait(bar(BAR_T_READY), (uint32_t)t & 1u);
              tc_fence_after();
              TC_STAMP(3);   // MMA: t operand ready seen
            }
            b = t_lo0;
#pragma unroll 1
            for (int c = 0; c < ku_full; ++c) {
              chunk64(16384u, d, b, 1u);
              b += kK64;
            }
            if (ku_rem) chunk(ku_rem, (uint32_t)ku_rem * 4096u, desc_hi((uint32_t)ku_rem * 256u), d, b, 1u);
            if (kS2Bufs == 1 && g == 1) umma_commit(bar(BAR_S2_FULL0), elected);   // half-buffer A (tiles i, g) complete
            d += (uint32_t)NS;
          }
        } else {
          // early pass of this unit block
          if (ke_one > 0) {
            gates_block(ke_one, d0, be0, 0u);
          } else {
            uint32_t d = d0;
#pragma unroll 1
            for (int g = 0; g < 4; ++g) {
              uint32_t b = be0;
              uint32_t acc = 0u;
#pragma unroll 1
              for (int c = 0; c < ke_full; ++c) {
                chunk64(16384u, d, b, acc);
                acc = 1u;
                b += kK64;
              }
              if (ke_rem) chunk(ke_rem, (uint32_t)ke_rem * 4096u, desc_hi((uint32_t)ke_rem * 256u), d, b, acc);
              d += (uint32_t)NS;
            }
          }
          if (ub == nub - 1 && !p.has_s1w) {   // layer 0: x(t) is consumed by the early passes
            umma_commit(bar(BAR_IN_EMPTY + in_s), elected);
            if (++in_s == nst) { in_s = 0; in_ph ^= 1u; }
          }
          if (ub == 0) {   // ---- the recurrent part needs t_u(t)
            mbar_wait(bar(BAR_T_READY), (uint32_t)t & 1u);
            tc_fence_after();
            TC_STAMP(3);   // MMA: t operand ready seen
          }
          // late pass: z += A2u . t_u(t)
          if (ku_one > 0) {
            gates_block(ku_one, d0, t_lo0, 1u);
          } else {
            uint32_t d = d0;
#pragma unroll 1
            for (int g = 0; g < 4; ++g) {
              uint32_t b = t_lo0;
#pragma unroll 1
              for (int c = 0; c < ku_full; ++c) {
                chunk64(16384u, d, b, 1u);
                b += kK64;
              }
              if (ku_rem) chunk(ku_rem, (uint32_t)ku_rem * 4096u, desc_hi((uint32_t)ku_rem * 256u), d, b, 1u);
              d += (uint32_t)NS;
            }
          }
        }
        if (kS2Bufs == 1) {
          if (!streaming) umma_commit(bar(BAR_S2_FULL0), elected);
          umma_commit(bar(BAR_S2_FULL1), elected);   // half-buffer B (and, resident, A) complete
        } else {
          umma_commit(bar(BAR_S2_FULL0 + buf), elected);
        }
        TC_STAMP(4 + (ub & 1));   // MMA: S2 block ub issued + committed
      }
      s2_use += (uint32_t)nub;
      // ---- S1w of the NEXT step: no recurrence, so it fills the tensor pipe while the epilogue works on z(t)
      if (p.has_s1w && t + 1 < T) issue_s1w();
    }
  } else if (warp >= 4) {
    // ======================= epilogue warps (256 threads; thread = TMEM lane = one row, CPT = NS/2 of the columns) ==
    const int ew = warp - 4;
    const int q = ew & 3;                      // TMEM lane quarter this warp may access (== warp % 4)
    const int half = ew >> 2;                  // column half
    const int c0 = half * CPT;
    const int row = q * 32 + lane;             // row of every 128-row tile handled by this thread
    const uint32_t lane_addr = ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
    float cst[NUB][CPT];                       // cell state of unit (ub*128+row), CPT sequences, FP32, all T steps
#pragma unroll
    for (int u = 0; u < NUB; ++u)
#pragma unroll
      for (int n = 0; n < CPT; ++n) cst[u][n] = 0.f;
    float bi[NUB][4];
#pragma unroll
    for (int u = 0; u < NUB; ++u)
#pragma unroll
      for (int g = 0; g < 4; ++g) bi[u][g] = p.bias[(u * 4 + g) * 128 + row];
    const int b_first = cta * NS + c0;         // global sequence index of this thread's first column
    if (STATE && p.c0 != nullptr) {
#pragma unroll
      for (int u = 0; u < NUB; ++u)
#pragma unroll
        for (int n = 0; n < CPT; ++n)
          if (b_first + n < p.B && u * 128 + row < p.Hu) cst[u][n] = p.c0[(size_t)(b_first + n) * p.Hu + u * 128 + row];
    }
    // ---- epilogue-1 plan of this thread, one bit per 128-row tile (everything the time loop would otherwise re-derive) ----
    const int e1_rows_u = p.ru + p.n_dense;    // rows of S1u anyone needs
    uint32_t u_ld = 0, u_st = 0, u_live = 0, w_ld = 0, w_st = 0, w_live = 0, x_ld = 0, x_st = 0, x_live = 0;
    int e1u_tiles = 0, e1w_tiles = 0;
    for (int mt = 0; mt < 4; ++mt) {   // up to 4 S1u tiles (layers without an S1w of their own), 2 S1w tiles
      const int r0 = mt * 128, j = r0 + row;
      if (p.store_x) {
        const int lo = r0 + q * 32;   // this warp's 32 rows of the tile
        if (r0 < p.rows_u && lo + 32 > p.x_r0 && lo < p.x_r0 + p.rx_pad) x_ld |= 1u << mt;   // warp-uniform
        if (j >= p.x_r0 && j < p.x_r0 + p.rx_pad) x_st |= 1u << mt;
        if (j >= p.x_r0 && j < p.x_r0 + p.rx) x_live |= 1u << mt;
      }
      if (r0 < p.rows_u && r0 + q * 32 < e1_rows_u) { u_ld |= 1u << mt; e1u_tiles = mt + 1; }   // warp-uniform
      if (j < p.ru_pad) u_st |= 1u << mt;
      if (j < p.ru) u_live |= 1u << mt;
      if (p.has_s1w && r0 < p.rows_w && r0 + q * 32 < p.rw_pad) { w_ld |= 1u << mt; e1w_tiles = mt + 1; }
      if (j < p.rw_pad) w_st |= 1u << mt;
      if (j < p.rw) w_live |= 1u << mt;
    }
    const uint32_t tb_u = sbase + sp.tbuf + act_offset(row, c0, NS);              // + kRowBlk per 128-row tile
    const uint32_t tb_w = tb_u + (uint32_t)(p.ru_pad >> 3) * (uint32_t)NS * 16u;
    const uint32_t tm_u = tm_s1u + lane_addr, tm_w = tm_s1w + lane_addr;          // + NS columns per tile
    const uint32_t hb_addr = sbase + sp.hbuf + act_offset(row, c0, NS);           // + kRowBlk per 128-unit block
    // Dense-top row owned by this thread (if any): its S1u row index is ru + o
    int y_mt = -1, y_o = 0;
    float y_bias = 0.f;
    for (int o = 0; o < p.n_dense; ++o)
      if (((p.ru + o) & 127) == row) { y_mt = (p.ru + o) >> 7; y_o = o; y_bias = p.dense_bias[o]; }
    float* y_ptr = (p.n_dense > 0 && y_mt >= 0) ? p.y + (size_t)b_first * T * p.n_dense + y_o : nullptr;   // + t*n_dense, + n*T*n_dense
    const int y_valid = p.B - b_first < CPT ? (p.B - b_first < 0 ? 0 : p.B - b_first) : CPT;
    const size_t y_seq_stride = (size_t)T * p.n_dense;
    // CPT fp32 accumulator columns of one row -> f16, 16-byte stores (8 sequences each) into an MN-major activation tile
    auto store_row = [&](uint32_t saddr, const uint32_t* r, bool live) {
      const uint32_t m = live ? 0xFFFFFFFFu : 0u;   // padding rows (rank..rank_pad) must hold zeros
#pragma unroll
      for (int j8 = 0; j8 < CPT / 8; ++j8)
        sts128(saddr + 128u * (uint32_t)j8, pack_f16(__uint_as_float(r[8 * j8 + 0]), __uint_as_float(r[8 * j8 + 1])) & m,
               pack_f16(__uint_as_float(r[8 * j8 + 2]), __uint_as_float(r[8 * j8 + 3])) & m,
               pack_f16(__uint_as_float(r[8 * j8 + 4]), __uint_as_float(r[8 * j8 + 5])) & m,
               pack_f16(__uint_as_float(r[8 * j8 + 6]), __uint_as_float(r[8 * j8 + 7])) & m);
    };
    auto load_cols = [&](uint32_t taddr, uint32_t* r) {   // CPT consecutive accumulator columns of this thread's TMEM lane
#pragma unroll
      for (int j16 = 0; j16 < CPT / 16; ++j16) tmem_ld16(taddr + 16u * (uint32_t)j16, r + 16 * j16);
      if constexpr (CPT == 8) tmem_ld8(taddr, r);
    };
    // t_w(tt) accumulators -> f16 rows [ru_pad, ru_pad + rw_pad) of the S2 B operand.  Runs one step ahead of its use
    // (S1w has no recurrence), at the tail of the previous step's epilogue, so it is never on the critical path.
    auto e1w = [&](uint32_t parity) {
      mbar_wait(bar(BAR_S1W_FULL), parity);
      tc_fence_after();
      uint32_t aw[CPT];
#pragma unroll 1
      for (int mt = 0; mt < e1w_tiles; ++mt) {
        load_cols(tm_w + (uint32_t)(mt * NS), aw);
        tmem_ld_wait();
        if ((w_st >> mt) & 1u) store_row(tb_w + (uint32_t)mt * kRowBlk, aw, (w_live >> mt) & 1u);
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(bar(BAR_TW_READY));
    };
    if (p.has_s1w) e1w(0u);
    uint32_t s2_use = 0;
    for (int t = 0; t < n_steps; ++t) {
      // ---- epilogue 1: t_u accumulators -> f16 rows of the S2 B operand; Dense-top row -> y(t-1) ----
      mbar_wait(bar(BAR_S1_FULL), (uint32_t)t & 1u);
      tc_fence_after();
      if (threadIdx.x == 128) TC_STAMP(8);   // EPI: S1 accumulators seen
      float yv[CPT];
      {
        uint32_t au[CPT];
        const bool step_live = t < T;   // the flush pass (t == T) only extracts the Dense-top row
#pragma unroll 1
        for (int mt = 0; mt < e1u_tiles; ++mt) {
          load_cols(tm_u + (uint32_t)(mt * NS), au);
          tmem_ld_wait();
          if (threadIdx.x == 128) TC_STAMP(0);   // EPI: S1 accumulator tile in registers
          if (((u_st >> mt) & 1u) && step_live) store_row(tb_u + (uint32_t)mt * kRowBlk, au, (u_live >> mt) & 1u);
          if (mt == y_mt) {   // keep the Dense-top row; it is written out after the t operand has been published
#pragma unroll
            for (int n = 0; n < CPT; ++n) yv[n] = __uint_as_float(au[n]) + y_bias;
          }
        }
      }
      if (t < T) {
        if (threadIdx.x == 128) TC_STAMP(1);   // EPI: t rows stored
        tc_fence_before();
        fence_proxy_async();
        if (threadIdx.x == 128) TC_STAMP(6);   // EPI: proxy fence done
        mbar_arrive(bar(BAR_T_READY));
        if (threadIdx.x == 128) TC_STAMP(9);   // EPI: t operand written + arrived
      }
      if (p.store_x && t > 0) {   // the next layer's t_w(t-1): off the critical path, after the t operand has been published
        if (t > 1) mbar_wait(bar(BAR_H_STORED), (uint32_t)(t - 2) & 1u);   // the tile of step t-2 has been shipped
        uint32_t ax[CPT];
#pragma unroll 1
        for (int mt = 0; mt < 4; ++mt) {
          if (!((x_ld >> mt) & 1u)) continue;
          load_cols(tm_u + (uint32_t)(mt * NS), ax);
          tmem_ld_wait();
          if ((x_st >> mt) & 1u) store_row(sbase + sp.xbuf + act_offset(mt * 128 + row - p.x_r0, c0, NS), ax, (x_live >> mt) & 1u);
        }
        fence_proxy_async();
        mbar_arrive(bar(BAR_H_DONE));
      }
      if (t > 0 && y_ptr != nullptr) {   // Dense top of step t-1 (off the critical path: after the arrive)
        float* yp = y_ptr + (size_t)(t - 1) * p.n_dense;
#pragma unroll
        for (int n = 0; n < CPT; ++n)
          if (n < y_valid) yp[(size_t)n * y_seq_stride] = yv[n];
      }
      if (t == T) break;

      // ---- epilogue 2: gates + cell update per unit block ------------------------------------------
      // the h(t-1) tile must have been copied out before it is overwritten
      if (p.store_h && t > 0) mbar_wait(bar(BAR_H_STORED), (uint32_t)(t - 1) & 1u);
#pragma unroll
      for (int ub = 0; ub < NUB; ++ub) {
        const uint32_t use = s2_use + (uint32_t)ub;
        const uint32_t buf = kS2Bufs == 2 ? (use & 1u) : 0u, turn = kS2Bufs == 2 ? (use >> 1) : use;
        mbar_wait(bar(BAR_S2_FULL0 + buf), turn & 1u);   // NS=64: half-buffer A = tiles (i, g)
        tc_fence_after();
        if (threadIdx.x == 128) TC_STAMP(10 + 2 * (ub & 1));   // EPI: z block ub seen
        const uint32_t tb = tm_s2 + buf * (uint32_t)(4 * NS) + lane_addr;   // tile slots: i, g(cell), f, o
        // staged: (i, g) -> the input product i*g, then (f, o), which frees half-buffer A early.  Inside each stage the
        // accumulators come out of TMEM in chunks of 8 columns, the loads of chunk k+1 in flight while chunk k goes through
        // the MUFU: TMEM reads (64 B/clk/SM: 4 096 cycles for the 256 KB of gate accumulators of one step) and the 5 tanh per
        // cell (5 120 cycles) are the two floors of this epilogue and used to run back to back, not overlapped.
        float ig[CPT];
        {
          uint32_t za[2][8], zb[2][8];
          tmem_ld8(tb + 0u * (uint32_t)NS, za[0]);
          tmem_ld8(tb + 1u * (uint32_t)NS, zb[0]);
#pragma unroll
          for (int k = 0; k < CPT / 8; ++k) {
            tmem_ld_wait();
            if (k + 1 < CPT / 8) {
              tmem_ld8(tb + 0u * (uint32_t)NS + 8u * (uint32_t)(k + 1), za[(k + 1) & 1]);
              tmem_ld8(tb + 1u * (uint32_t)NS + 8u * (uint32_t)(k + 1), zb[(k + 1) & 1]);
            } else if (kS2Bufs == 1) {   // every load of half-buffer A has completed
              tc_fence_before();
              mbar_arrive(bar(BAR_S2_EMPTY0));
            }
#pragma unroll
            for (int m = 0; m < 8; ++m)
              ig[8 * k + m] = fmaf(0.5f, tanh_approx(__uint_as_float(za[k & 1][m]) + bi[ub][0]), 0.5f) *
                              tanh_approx(__uint_as_float(zb[k & 1][m]) + bi[ub][1]);
          }
        }
        if (kS2Bufs == 1) {
          mbar_wait(bar(BAR_S2_FULL1), turn & 1u);         // half-buffer B = tiles (f, o)
          tc_fence_after();
        }
        {
          uint32_t zf[2][8], zo[2][8];
          tmem_ld8(tb + 2u * (uint32_t)NS, zf[0]);
          tmem_ld8(tb + 3u * (uint32_t)NS, zo[0]);
#pragma unroll
          for (int k = 0; k < CPT / 8; ++k) {
            tmem_ld_wait();
            if (k + 1 < CPT / 8) {
              tmem_ld8(tb + 2u * (uint32_t)NS + 8u * (uint32_t)(k + 1), zf[(k + 1) & 1]);
              tmem_ld8(tb + 3u * (uint32_t)NS + 8u * (uint32_t)(k + 1), zo[(k + 1) & 1]);
            } else {   // the accumulators are in registers: hand the TMEM buffer back to the MMA warp right away
              tc_fence_before();
              mbar_arrive(bar(kS2Bufs == 1 ? BAR_S2_EMPTY1 : BAR_S2_EMPTY0 + buf));
            }
            float hv[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) {
              const int n = 8 * k + m;
              const float fg = fmaf(0.5f, tanh_approx(__uint_as_float(zf[k & 1][m]) + bi[ub][2]), 0.5f);
              const float og = fmaf(0.5f, tanh_approx(__uint_as_float(zo[k & 1][m]) + bi[ub][3]), 0.5f);
              const float c = fmaf(fg, cst[ub][n], ig[n]);
              cst[ub][n] = c;
              hv[m] = og * tanh_approx(c);
            }
            sts128(hb_addr + (uint32_t)ub * kRowBlk + 128u * (uint32_t)k, pack_f16(hv[0], hv[1]), pack_f16(hv[2], hv[3]),
                   pack_f16(hv[4], hv[5]), pack_f16(hv[6], hv[7]));
          }
        }
        fence_proxy_async();
        mbar_arrive(bar(BAR_H_READY + ub));
        if (threadIdx.x == 128) TC_STAMP(11 + 2 * (ub & 1));   // EPI: block ub done
      }
      s2_use += (uint32_t)nub;
      if (p.store_h) mbar_arrive(bar(BAR_H_DONE));
      if (p.has_s1w && t + 1 < T) e1w((uint32_t)(t + 1) & 1u);
      if (threadIdx.x == 128) TC_STAMP(14);      // EPI: h(t) published
    }
    if (STATE && p.h_n != nullptr) {   // return_state: h(T-1) as this engine holds it (FP16 between steps): read back this thread's own stores
      for (int u = 0; u < NUB; ++u)
        for (int n = 0; n < CPT; ++n)
          if (b_first + n < p.B && u * 128 + row < p.Hu)
            p.h_n[(size_t)(b_first + n) * p.Hu + u * 128 + row] =
                __half2float(*reinterpret_cast<const __half*>(smem + sp.hbuf + act_offset(u * 128 + row, c0 + n, NS)));
    }
    if (STATE && p.c_n != nullptr) {   // return_state: c(T-1), float32 all along
#pragma unroll
      for (int u = 0; u < NUB; ++u)
#pragma unroll
        for (int n = 0; n < CPT; ++n)
          if (b_first + n < p.B && u * 128 + row < p.Hu) p.c_n[(size_t)(b_first + n) * p.Hu + u * 128 + row] = cst[u][n];
    }
  }

  // ---- teardown -----------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
  (void)stamp_cta;
}

// layers launched one after the other (any batch size): grid = tiles of one layer
template <int NUB, bool STREAM, int NS, int EW = kEpiWarps, bool STATE = false>
__global__ void __launch_bounds__(32 * (4 + EW), 1) lstm_tc_layer_kernel(const TcLayerParams p) {
  tc_layer_body<NUB, STREAM, NS, EW, STATE>(p, (int)blockIdx.x, nullptr, nullptr, blockIdx.x == 0);
}

// ALL layers in one (cooperative, fully co-resident) launch: CTA b runs layer b / n_tiles on tile b % n_tiles and the layers
// form a pipeline over time -- layer l+1 consumes h_l(t) a few steps after layer l produced it.  With L layers x B/NS tiles
// <= the SM count this doubles the sequences per SM (NS = 64) without idling SMs, which halves the shared-memory-port cost
// per sequence of re-streamed weights.
struct TcPipeParams {
  TcLayerParams layer[kMaxLayers];
  int n_layers, n_tiles;
  int* progress;   // [n_layers][n_tiles] steps published, zeroed before the launch
  int* consumed;   // [n_layers][n_tiles] steps of its INPUT hand-off ring a layer has finished reading, zeroed before the launch
};
template <int NUB, int NS, int EW = kEpiWarps, bool RAWX = false>
__global__ void __launch_bounds__(32 * (4 + EW), 1) lstm_tc_pipe_kernel(const __grid_constant__ TcPipeParams pp) {
  const int layer = (int)blockIdx.x / pp.n_tiles, tile = (int)blockIdx.x - layer * pp.n_tiles;
  const TcLayerParams& p = pp.layer[layer];
  const int* prog_in = layer > 0 ? pp.progress + (size_t)(layer - 1) * pp.n_tiles + tile : nullptr;
  int* prog_out = layer + 1 < pp.n_layers ? pp.progress + (size_t)layer * pp.n_tiles + tile : nullptr;
  int* cons_pub = layer > 0 ? pp.consumed + (size_t)layer * pp.n_tiles + tile : nullptr;
  const int* cons_wait = layer + 1 < pp.n_layers ? pp.consumed + (size_t)(layer + 1) * pp.n_tiles + tile : nullptr;
  if (p.streaming) tc_layer_body<NUB, true, NS, EW, false, RAWX>(p, tile, prog_in, prog_out, tile == 0, cons_pub, cons_wait);
  else if constexpr (NUB <= 4 && !RAWX) tc_layer_body<NUB, false, NS, EW>(p, tile, prog_in, prog_out, tile == 0, cons_pub, cons_wait);   // (H > 512 never fits resident; RAWX is a streamed-only build)
}

// ------------------------------------------------------------------------------------------------
// packing / layout kernels (run once per weight update, or once per forward for the sequences)
// ------------------------------------------------------------------------------------------------

// effective right-factor element of a block: row kk of the (rank x 4H) matrix, gate column n
__device__ __forceinline__ float block_right(const Block& b, int kk, int n) {
  int rel = n - b.out0;
  if (rel < 0) return 0.f;
  if (b.ident) {
    if (rel < b.rank) return rel == kk ? 1.f : 0.f;
    rel -= b.rank;
  }
  if (rel >= b.ncols) return 0.f;
  return b.right[(size_t)kk * b.right_ld + rel];
}
__device__ __forceinline__ float block_left(const Block& b, int k, int j) {   // (L sigma)[k][j]; a full (unfactored) matrix is I . W
  const float l = b.left ? b.left[(size_t)k * b.left_ld + j] : (k == j ? 1.f : 0.f);
  return l * (b.scale ? b.scale[j] : 1.f);
}

// H = padded units (multiple of 128: the kernel's row-tile granularity), Hu = the layer's true units, D = its true input width.
// Padded cells have zero weights and zero bias: z = 0 -> i = f = o = 1/2, g = 0 -> c = h = 0 for all t, exactly.
__global__ void pack_wstream_kernel(const PackChunk* __restrict__ chunks, Block bw, Block bu, Block bw_next, int x_r0, int rx, int H, int Hu, int D,
                                    int ru_pad, const float* __restrict__ dense_k, int n_dense, int n_out, __half* __restrict__ img) {
  const PackChunk c = chunks[blockIdx.x];
  __half* out = img + c.byte_off / 2;
  const int total = c.rows * c.kc;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int row = idx / c.kc, k = idx - row * c.kc;
    float v = 0.f;
    if (c.seg == 0) {
      const int j = c.r0 + row, kk = c.k0 + k;
      if (j < bw.rank && kk < D) v = block_left(bw, kk, j);
    } else if (c.seg == 1) {
      const int j = c.r0 + row, kk = c.k0 + k;
      if (kk < Hu) {
        if (j < bu.rank) v = block_left(bu, kk, j);
        else if (j - bu.rank < n_dense) v = dense_k[(size_t)kk * n_out + (j - bu.rank)];
        else if (rx > 0 && j >= x_r0 && j - x_r0 < rx) v = block_left(bw_next, kk, j - x_r0);   // the next layer's (L_w sigma_w)^T rows
      }
    } else {
      const int gate = (c.g == 1) ? 2 : (c.g == 2) ? 1 : c.g;   // tile slots are ordered i, g(cell), f, o; Keras columns i, f, c, o
      const int unit = c.ub * 128 + row;
      const int n = gate * Hu + unit;
      const int kk = c.k0 + k;
      if (unit < Hu) {
        if (c.part == 0) {            // recurrent part: row kk of R_u
          if (kk < bu.rank) v = block_right(bu, kk, n);
        } else if (c.part == 2) {     // input part, layers >= 1: row kk of R_w
          if (kk < bw.rank) v = block_right(bw, kk, n);
        } else if (kk < D) {          // input part, layer 0: the dense D x 4H product W0 = (L_w sigma_w) R_w
          float acc = 0.f;
          for (int j = 0; j < bw.rank; ++j) acc = fmaf(block_left(bw, kk, j), block_right(bw, j, n), acc);
          v = acc;
        }
      }
      if (gate != 2) v *= 0.5f;   // sigmoid gates: tanh(z/2) form
    }
    const size_t off = (size_t)(row / 8) * ((size_t)c.kc * 16) + (size_t)(k / 8) * 128 + (size_t)(row % 8) * 16 + (size_t)(k % 8) * 2;
    out[off / 2] = __float2half_rn(v);
  }
}

__global__ void pack_bias_kernel(const float* __restrict__ bias, int H, int Hu, float* __restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // [ub][slot][128], slots ordered i, g(cell), f, o
  if (idx >= 4 * H) return;
  const int ub = idx / 512, slot = (idx / 128) & 3, i = idx & 127;
  const int g = (slot == 1) ? 2 : (slot == 2) ? 1 : slot;
  const int unit = ub * 128 + i;
  img[idx] = unit < Hu ? bias[g * Hu + unit] * (g != 2 ? 0.5f : 1.f) : 0.f;
}

// x (B,T,D) fp32 -> f16 activation tile images [cta][t] (K = Dpad16).  One block = one batch tile x TT steps:
// coalesced reads of each sequence's TT*D contiguous floats into smem, then contiguous 16-byte tile writes.
__global__ void __launch_bounds__(256) pack_x_kernel(const float* __restrict__ x, int B, int T, int D, int Dpad, int TT, int ns,
                                                     uint8_t* __restrict__ img) {
  extern __shared__ float xs[];   // [ns][TT*D]
  const uint32_t tile = act_tile_bytes(Dpad, ns);
  const int cta = blockIdx.y;
  const int t0 = blockIdx.x * TT;
  const int nt = imin(TT, T - t0);
  const int row = nt * D;   // contiguous floats per sequence
  for (int idx = threadIdx.x; idx < ns * row; idx += blockDim.x) {
    const int n = idx / row, o = idx - n * row;
    const int b = cta * ns + n;
    xs[n * (TT * D) + o] = (b < B) ? x[((size_t)b * T + t0) * D + o] : 0.f;
  }
  __syncthreads();
  // one 16-byte store = 8 consecutive sequences of one (t, k)
  const int ngr = ns / 8;                 // 16-byte words per (t, k): one per group of 8 sequences
  const int per_t = Dpad * ngr;
  for (int idx = threadIdx.x; idx < nt * per_t; idx += blockDim.x) {
    const int tt = idx / per_t, rem = idx - tt * per_t;
    const int kg = rem / (8 * ngr), w = rem - kg * (8 * ngr);   // within a k-group of 8: [n-group][k%8] 16-byte words
    const int ng = w / 8, k = kg * 8 + (w & 7);
    uint32_t v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = ng * 8 + 2 * i;
      const float a = k < D ? xs[n * (TT * D) + tt * D + k] : 0.f;
      const float c = k < D ? xs[(n + 1) * (TT * D) + tt * D + k] : 0.f;
      v[i] = pack_f16(a, c);
    }
    *reinterpret_cast<uint4*>(img + ((size_t)cta * T + t0 + tt) * tile + act_offset(k, ng * 8, ns)) = make_uint4(v[0], v[1], v[2], v[3]);
  }
}

// last layer h tiles -> y (B,T,H) fp32 (only when the model has no Dense top; the Dense top is fused otherwise)
__global__ void unpack_out_kernel(const uint8_t* __restrict__ img, int B, int T, int Hpad, int H, int ns, float* __restrict__ y) {
  const uint32_t tile = act_tile_bytes(Hpad, ns);
  const int cta = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int item = blockIdx.x * nw + warp; item < T * ns; item += gridDim.x * nw) {
    const int t = item / ns, n = item - t * ns;
    const int b = cta * ns + n;
    if (b >= B) continue;
    const uint8_t* tp = img + ((size_t)cta * T + t) * tile;
    for (int k = lane; k < H; k += 32) y[((size_t)b * T + t) * H + k] = __half2float(*reinterpret_cast<const __half*>(tp + act_offset(k, n, ns)));
  }
}

// Dense / TimeDistributed(Dense) top from the last hidden-sequence image (only when it could not be fused into S1u):
// y[b, t, o] = sum_k h[k][b] W[k][o] + bias[o].  One block per (step, tile): the 16-byte words of the tile (8 sequences of
// one unit each) are read coalesced; thread (q, n-group) accumulates a quarter of the units for 8 sequences.
__global__ void __launch_bounds__(256) dense_top_kernel(const uint8_t* __restrict__ img, int B, int T, int Hpad, int H, int ns,
                                                        const float* __restrict__ dk, const float* __restrict__ db, int n_out,
                                                        float* __restrict__ y) {
  __shared__ float part[64][8][8];   // [k-slice (<= 64)][n-group][sequence in group]
  const uint32_t tile = act_tile_bytes(Hpad, ns);
  const int cta = blockIdx.y, t = blockIdx.x;
  const int ngr = ns / 8, nks = 256 / ngr;          // n-groups, k-slices
  const int ng = threadIdx.x % ngr, ks = threadIdx.x / ngr;
  const uint8_t* tp = img + ((size_t)cta * T + t) * tile;
  for (int o = 0; o < n_out; ++o) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = ks; k < H; k += nks) {
      const uint4 v = *reinterpret_cast<const uint4*>(tp + act_offset(k, ng * 8, ns));
      const float w = dk[(size_t)k * n_out + o];
      const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h2[i]);
        acc[2 * i] = fmaf(f.x, w, acc[2 * i]);
        acc[2 * i + 1] = fmaf(f.y, w, acc[2 * i + 1]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) part[ks][ng][i] = acc[i];
    __syncthreads();
    if ((int)threadIdx.x < ns) {
      const int n = threadIdx.x, b = cta * ns + n;
      float sum = db[o];
      for (int s2 = 0; s2 < nks; ++s2) sum += part[s2][n / 8][n % 8];
      if (b < B) y[((size_t)b * T + t) * n_out + o] = sum;
    }
    __syncthreads();
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct TcLayerImg {
  PackChunk* chunks = nullptr;
  uint32_t* slots = nullptr;
  uint8_t* wimg = nullptr;
  float* bias = nullptr;
  TcLayerParams prm;
};

struct TcState {
  TcLayerImg layers[kMaxLayers];
  int n_layers = 0;
  int ns = 0;   // tile width the images / plans were built for (the chunk order depends on resident vs streamed)
};

// Per-device scratch shared by every handle: the FP16 image of x and the two ping-pong hidden-sequence images
// (C3: 134 MB + 2 x 2.1 GB).  A rank sweep holds hundreds of handles; per-handle scratch would not fit.  Work on
// one stream is ordered, so sharing is safe there; a forward on a different stream first drains the previous one.
struct TcWorkspace {
  uint8_t* seq[kMaxLayers] = {};     // hidden-sequence images: slots l & 1 (layers in turn) or l (layers pipelined)
  size_t seq_bytes[kMaxLayers] = {};
  int* progress = nullptr;           // layer-pipelined launch: [layer][tile] published-step counters
  size_t progress_elems = 0;
  uint8_t* xseq = nullptr;
  size_t xseq_bytes = 0;
  // Ordering across streams: every forward records `done` after its last launch and the next forward makes ITS stream wait on
  // it (no stream handle is remembered -- the caller may have destroyed it).  `mu` serialises host threads sharing the device.
  cudaEvent_t done = nullptr;
  std::mutex mu;
};
static TcWorkspace g_tc_ws[16];

// ---- 2-factor (ReducedLSTMCell) blocks on the tensor cores ---------------------------------------------------------------
// z = [a | a C] with a = in . B  (svd_classes_v3.py:321-328).  C = inv(V1) V2 has entries in the hundreds on trained weights
// (cond(V1) ~ 1e2): FP16 is the wrong container for it (5 % errors measured).  The SAME linear map is in . B' . Q^T with
// [I | C] = P diag(s) Q^T (thin SVD, K2 on device), B' = B P diag(s): Q^T has orthonormal rows (entries <= 1) and B' is
// as benign as the 3-factor (L sigma), so the block packs and runs exactly like a 3-factor one, at 3-factor accuracy.
__global__ void tc_ident_right_kernel(Block b, int n, float* __restrict__ R) {   // R (rank x n) = [I | C]
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (size_t)b.rank * n; idx += stride) {
    const int kk = (int)(idx / n), c = (int)(idx - (size_t)kk * n);
    R[idx] = block_right(b, kk, b.out0 + c);
  }
}
__global__ void tc_ident_left_kernel(Block b, int rows, const float* __restrict__ P, const float* __restrict__ S, float* __restrict__ out) {
  // out (rows x r) = left (rows x r) . P (r x r) . diag(S)
  const int r = b.rank;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (size_t)rows * r; idx += stride) {
    const int i = (int)(idx / r), j = (int)(idx - (size_t)i * r);
    double acc = 0.0;
    for (int k = 0; k < r; ++k) acc += (double)block_left(b, i, k) * (double)P[(size_t)k * r + j];
    out[idx] = (float)(acc * (double)S[j]);
  }
}

static int tc_reorthogonalise(const Block& b, int rows, cudaStream_t stream, Block* out, std::vector<void*>& temps, int* launches) {
  const int r = b.rank, n = b.rank + b.ncols;
  float *R = nullptr, *P = nullptr, *S = nullptr, *Qt = nullptr, *Bp = nullptr;
  SVD_CUDA_TRY(cudaMalloc(&R, sizeof(float) * r * n));
  temps.push_back(R);
  SVD_CUDA_TRY(cudaMalloc(&P, sizeof(float) * r * r));
  temps.push_back(P);
  SVD_CUDA_TRY(cudaMalloc(&S, sizeof(float) * r));
  temps.push_back(S);
  SVD_CUDA_TRY(cudaMalloc(&Qt, sizeof(float) * r * n));
  temps.push_back(Qt);
  SVD_CUDA_TRY(cudaMalloc(&Bp, sizeof(float) * rows * r));
  temps.push_back(Bp);
  tc_ident_right_kernel<<<imin(1184, (r * n + 255) / 256), 256, 0, stream>>>(b, n, R);
  if (int rc = svdlstm_svd_jacobi_batched(R, 1, r, n, P, S, Qt, nullptr, stream)) return rc;
  tc_ident_left_kernel<<<imin(1184, (rows * r + 127) / 128), 128, 0, stream>>>(b, rows, P, S, Bp);
  *launches += 3;
  *out = Block{Bp, nullptr, Qt, r, n, r, n, b.out0, 0, b.from_h, 0};
  return 0;
}

void tc_free(TcState* s) {
  if (!s) return;
  for (int l = 0; l < kMaxLayers; ++l) {
    if (s->layers[l].wimg) cudaFree(s->layers[l].wimg);
    if (s->layers[l].chunks) cudaFree(s->layers[l].chunks);
    if (s->layers[l].slots) cudaFree(s->layers[l].slots);
    if (s->layers[l].bias) cudaFree(s->layers[l].bias);
  }
  delete s;
}

// Which layers hand the NEXT layer's t_w over (instead of their own h)?  The extra S1u rows must fit the S1u TMEM tiles: 2 tiles of
// 128 rows at 64-sequence tiles, 3 at 32 -- or 4 when the layer has no S1w of its own (layer 0, and every layer that itself receives
// t_w: its S1w TMEM region is free).  Needs merged factored cells on both sides.  SVDLSTM_TC_HANDOFF=h disables it.
static int tc_s1u_tiles_max(int ns, bool no_s1w) { return no_s1w ? 4 : (ns == 64 ? 2 : 3); }
// ---- split cell forms (svd_classes_v3.py:146-232, 330-363: one factorisation per gate) ------------------------------------------
// The kernel consumes ONE (left, scale, right) triple per side (input / recurrent) of a layer.  A split layer's four per-gate
// blocks of a side are merged into such a triple when the weights are packed, by one of two routes:
//   concat  left = [L_i | L_f | L_c | L_o] (kin x R), scale likewise, right = the block-diagonal (R x 4H) matrix of the R_g:
//           the same contractions as the reference's split form (zero blocks are multiplied too); needs R = sum of the gate
//           ranks <= 256 and pays off when R < kin.  2-factor gate blocks ([I | C_g]) are re-orthogonalised one by one first
//           (tc_reorthogonalise), exactly like a merged 2-factor block;
//   dense   W = sum_g (L_g sigma_g) R_g placed at gate g's columns (float64 accumulation), run as I . W like an unfactored cell:
//           for R >= kin (a split model at full rank: R = 4 min(D, H)); needs kin <= 256.
// kTcRouteNone = neither fits.
enum { kTcRouteNone = 0, kTcRouteConcat = 1, kTcRouteDense = 2 };
static int tc_side_route(const LayerDesc& Ld, int side, int* rank_out) {
  const int kin = side == 0 ? Ld.d_in : Ld.units;
  int R = 0;
  for (int bi = 0; bi < Ld.n_blocks; ++bi)
    if (Ld.blocks[bi].from_h == side) R += Ld.blocks[bi].rank;
  if (R <= 256 && R < kin) { *rank_out = R; return kTcRouteConcat; }
  if (kin <= 256) { *rank_out = kin; return kTcRouteDense; }
  *rank_out = R;
  return kTcRouteNone;
}

// Shape-level merged equivalent of `md`: merged layers are copied, split layers become two blocks (pointers null: the device
// tensors are built by tc_merge_side when the weights are packed).  False if a split layer has no route.
static bool tc_effective_desc(const ModelDesc& md, ModelDesc& emd, const char** why) {
  emd = md;
  for (int l = 0; l < md.n_layers; ++l) {
    const LayerDesc& Ld = md.layers[l];
    if (Ld.n_blocks == 2) continue;
    LayerDesc& E = emd.layers[l];
    E.n_blocks = 2;
    int p = 0;
    for (int side = 0; side < 2; ++side) {
      int rank = 0;
      if (tc_side_route(Ld, side, &rank) == kTcRouteNone) {
        *why = "split cell form: the gate ranks of one side sum to more than 256 and its input is wider than 256";
        return false;
      }
      Block b{};
      b.rank = rank;
      b.ncols = 4 * Ld.units;
      b.right_ld = 4 * Ld.units;
      b.left_ld = rank;
      b.from_h = side;
      b.p_off = p;
      p += rank;
      E.blocks[side] = b;
    }
    E.p_total = p;
  }
  return true;
}

struct TcSideBlocks {
  Block b[4];
  int n;
};
// concat route: left_cat (kin x R), scale_cat (R), right_bd (R x 4H, zero outside the gates' own columns)
__global__ void tc_concat_side_kernel(TcSideBlocks sb, int kin, int R, int H4, float* __restrict__ left_cat, float* __restrict__ scale_cat,
                                      float* __restrict__ right_bd) {
  const int n_left = kin * R, n_right = R * H4;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_left + R + n_right; idx += gridDim.x * blockDim.x) {
    int p0 = 0;
    if (idx < n_left + R) {
      const bool is_scale = idx >= n_left;
      const int i = is_scale ? 0 : idx / R, k = is_scale ? idx - n_left : idx - i * R;
      float v = 0.f;
      for (int q = 0; q < sb.n; ++q) {
        const Block& b = sb.b[q];
        if (k >= p0 && k < p0 + b.rank) {
          const int kk = k - p0;
          v = is_scale ? (b.scale ? b.scale[kk] : 1.f) : (b.left ? b.left[(size_t)i * b.left_ld + kk] : (i == kk ? 1.f : 0.f));
        }
        p0 += b.rank;
      }
      if (is_scale) scale_cat[k] = v;
      else left_cat[idx] = v;
    } else {
      const int r = idx - n_left - R;
      const int k = r / H4, n = r - k * H4;
      float v = 0.f;
      for (int q = 0; q < sb.n; ++q) {
        const Block& b = sb.b[q];
        if (k >= p0 && k < p0 + b.rank) {
          const int rel = n - b.out0;
          if (rel >= 0 && rel < b.ncols) v = b.right[(size_t)(k - p0) * b.right_ld + rel];
        }
        p0 += b.rank;
      }
      right_bd[r] = v;
    }
  }
}
// dense route: W (kin x 4H) = sum over the side's blocks of (L sigma) [I |] R at the block's columns
__global__ void tc_dense_side_kernel(TcSideBlocks sb, int kin, int H4, float* __restrict__ w) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < kin * H4; idx += gridDim.x * blockDim.x) {
    const int i = idx / H4, n = idx - i * H4;
    double acc = 0.0;
    for (int q = 0; q < sb.n; ++q) {
      const Block& b = sb.b[q];
      int rel = n - b.out0;
      if (rel < 0) continue;
      if (b.ident) {
        if (rel < b.rank) {
          acc += (double)(b.left ? b.left[(size_t)i * b.left_ld + rel] : (i == rel ? 1.f : 0.f)) * (double)(b.scale ? b.scale[rel] : 1.f);
          continue;
        }
        rel -= b.rank;
      }
      if (rel >= b.ncols) continue;
      for (int kk = 0; kk < b.rank; ++kk) {
        const double lv = b.left ? (double)b.left[(size_t)i * b.left_ld + kk] : (i == kk ? 1.0 : 0.0);
        acc += lv * (double)(b.scale ? b.scale[kk] : 1.f) * (double)b.right[(size_t)kk * b.right_ld + rel];
      }
    }
    w[idx] = (float)acc;
  }
}
// One side of a split layer -> the merged block `out` (shape from tc_effective_desc) with device tensors in `temps`.
static int tc_merge_side(const LayerDesc& Ld, int side, cudaStream_t stream, Block* out, std::vector<void*>& temps, int* launches) {
  TcSideBlocks sb{};
  for (int bi = 0; bi < Ld.n_blocks; ++bi)
    if (Ld.blocks[bi].from_h == side) {
      SVD_REQUIRE(sb.n < 4, "tensor-core engine: more than four blocks on one side of a split layer");
      sb.b[sb.n++] = Ld.blocks[bi];
    }
  const int kin = side == 0 ? Ld.d_in : Ld.units, H4 = 4 * Ld.units;
  int rank = 0;
  const int route = tc_side_route(Ld, side, &rank);
  SVD_REQUIRE(route != kTcRouteNone && rank == out->rank, "tensor-core engine: internal error (split-form route)");
  auto alloc = [&](float** q, size_t n) -> int {
    SVD_CUDA_TRY(cudaMalloc(q, sizeof(float) * n));
    temps.push_back(*q);
    return 0;
  };
  if (route == kTcRouteConcat) {
    for (int q = 0; q < sb.n; ++q)
      if (sb.b[q].ident) {
        Block ortho;
        if (int rc = tc_reorthogonalise(sb.b[q], kin, stream, &ortho, temps, launches)) return rc;
        sb.b[q] = ortho;
      }
    float *lc = nullptr, *sc = nullptr, *rb = nullptr;
    if (int e = alloc(&lc, (size_t)kin * rank)) return e;
    if (int e = alloc(&sc, (size_t)rank)) return e;
    if (int e = alloc(&rb, (size_t)rank * H4)) return e;
    tc_concat_side_kernel<<<296, 256, 0, stream>>>(sb, kin, rank, H4, lc, sc, rb);
    out->left = lc;
    out->scale = sc;
    out->right = rb;
  } else {
    float* w = nullptr;
    if (int e = alloc(&w, (size_t)kin * H4)) return e;
    tc_dense_side_kernel<<<296, 256, 0, stream>>>(sb, kin, H4, w);
    out->left = nullptr;
    out->scale = nullptr;
    out->right = w;
  }
  SVD_CUDA_TRY(cudaGetLastError());
  ++*launches;
  return 0;
}

static void tc_handoff_chain(const ModelDesc& md, int ns, bool out[kMaxLayers]) {
  const char* e = getenv("SVDLSTM_TC_HANDOFF");
  const bool off = e && e[0] == 'h';
  for (int l = 0; l < kMaxLayers; ++l) out[l] = false;
  for (int l = 0; l + 1 < md.n_layers; ++l) {
    const LayerDesc& A = md.layers[l];
    const LayerDesc& B = md.layers[l + 1];
    if (off || A.n_blocks != 2 || B.n_blocks != 2) continue;
    const bool no_s1w = l == 0 || out[l - 1];
    out[l] = round_up(A.blocks[1].rank, 8) + round_up(B.blocks[0].rank, 8) <= 128 * tc_s1u_tiles_max(ns, no_s1w);
  }
}

static bool tc_layer_params(const ModelDesc& md, int l, int ns, TcLayerParams& p, const char** why) {
  const LayerDesc& L = md.layers[l];
  if (L.n_blocks != 2) { *why = "only merged (non-split) cell forms run on the tensor-core engine"; return false; }
  const Block& bw = L.blocks[0];
  const Block& bu = L.blocks[1];
  // (a full, unfactored cell runs as the "factorisation" I . W: its first contraction is an identity, exact in FP16 -- the baseline the
  //  truncated models are compared with on the same engine; it pays one wasted H x H contraction per step)
  // units are padded to the 128-row tile of the MMAs (zero weights and bias: padded cells stay at h = c = 0 exactly), which is
  // what lets the shipped 3 x 15 DROPBEAR model run on this engine too (its RMSE delta is part of north_star)
  const int Hu = L.units;
  const int H = round_up(Hu, 128);
  if (H > 128 * kMaxUB || (H > 512 && H != 1024)) { *why = "tensor-core engine needs units <= 512, or 513..1024 (run as 1024)"; return false; }
  if (H > 512 && ns != 32) { *why = "units above 512 run with 32-sequence tiles"; return false; }
  if (bu.rank > 256 || bw.rank > 256) { *why = "ranks above 256 are not supported by the tensor-core engine"; return false; }
  const bool last = (l == md.n_layers - 1);
  p = TcLayerParams{};
  p.ns = ns;
  p.H = H;
  p.Hu = Hu;
  p.ru = bu.rank;
  p.rw = bw.rank;
  p.ru_pad = round_up(bu.rank, 16);
  p.has_s1w = l > 0;
  // The Dense top rides in spare rows of the S1u tiles when they fit TMEM (3 tiles of 128 rows at 32-sequence tiles, 2 at
  // 64); otherwise the layer stores its hidden sequence and dense_top_kernel finishes the job.
  bool handoff[kMaxLayers];
  tc_handoff_chain(md, ns, handoff);
  const bool in_tw = l > 0 && handoff[l - 1], out_tw = handoff[l];
  const int s1_rows_max = 128 * tc_s1u_tiles_max(ns, l == 0 || in_tw);
  p.n_dense = (last && md.n_out > 0 && round_up(bu.rank + md.n_out, 8) <= s1_rows_max) ? md.n_out : 0;
  p.store_h = p.n_dense > 0 ? 0 : 1;
  p.rows_u = round_up(p.ru + p.n_dense, 8);
  if (out_tw) {
    p.store_x = 1;
    p.store_h = 0;
    p.rx = md.layers[l + 1].blocks[0].rank;
    p.rx_pad = round_up(p.rx, 16);
    p.x_r0 = p.rows_u;
    p.rows_u = round_up(p.x_r0 + p.rx, 8);
  }
  p.early_part = l == 0 ? 1 : 2;
  if (in_tw) p.has_s1w = 0;
  if (ns == 64 && H > 256) { *why = "64-sequence tiles support units <= 256"; return false; }
  if (l == 0) {
    if (L.d_in > 64) { *why = "layer-0 input_dim above 64 is not supported by the tensor-core engine yet"; return false; }
    p.Kin = round_up(L.d_in, 16);
    p.kx = p.Kin;
    p.rw_pad = 0;
    p.rows_w = 0;
  } else if (in_tw) {   // the input tiles ARE t_w(t): they enter the early part of S2 directly, like x(t) on layer 0
    p.Kin = round_up(bw.rank, 16);
    p.kx = p.Kin;
    p.rw_pad = 0;
    p.rows_w = 0;
  } else {
    p.Kin = round_up(md.layers[l - 1].units, 128);
    p.kx = 0;
    p.rw_pad = round_up(bw.rank, 16);
    p.rows_w = round_up(bw.rank, 8);
  }
  p.segw_bytes = p.segu_bytes = p.seg2_bytes = 0;
  p.n_chunks_w = p.n_chunks_u = p.n_chunks_2 = 0;
  if (p.has_s1w) for_seg_w(p, [&](uint32_t b, int, int, int) { p.segw_bytes += b; ++p.n_chunks_w; });
  for_seg_u(p, [&](uint32_t b, int, int, int, int) { p.segu_bytes += b; ++p.n_chunks_u; });
  for_seg_2(p, [&](uint32_t b, int, int, int, int, int) { p.seg2_bytes += b; ++p.n_chunks_2; });
  // resident if the whole stream fits next to the activation buffers, else a ring of 16 KB slots
  p.streaming = 0;
  p.w_slots = 1;
  p.in_stages = 3;
  if (tc_plan(p).total > kSmemCap) p.in_stages = 2;
  if (tc_plan(p).total > kSmemCap && p.has_s1w) p.in_stages = 1;   // the tile of step t+1 is fetched while step t computes
  // 64-sequence tiles always stream: with ONE S2 accumulator buffer the streamed issue order (tile by tile, half-buffer (i,g)
  // handed to the epilogue while (f,o) is still being issued) overlaps MMA and epilogue where the resident order (whole unit
  // block per asm block) cannot -- measured 4.95 vs 5.26 ms at ranks 8-16, 5.08 vs 5.47 ms at rank 32 on C3.
  if (tc_plan(p).total > kSmemCap || ns == 64) {
    p.streaming = 1;
    p.in_stages = p.has_s1w ? 1 : 2;   // every KB goes to the weight ring: its depth must cover the L2 latency
    // ring-slot fills: consecutive chunks of a segment share a slot up to slot_bytes (same greedy rule as the MMA warp's cursor).
    // 32 KB slots halve the hand-overs (commit + full wait) of the MMA warp, which is issue-bound; when fewer than four of them
    // fit next to the activation buffers (H = 1024) the ring falls back to 16 KB slots to keep enough copies in flight.
    auto count_slots = [&](auto&& for_seg, int& n) {
      uint32_t cur = p.slot_bytes + 1;   // forces a new slot at the segment start
      n = 0;
      for_seg([&](uint32_t b) {
        if (cur + b > p.slot_bytes) { ++n; cur = 0; }
        cur += b;
      });
    };
    for (uint32_t sb : {kSlotBytes, 16384u}) {
      p.slot_bytes = sb;
      p.n_slots_w = 0;
      if (p.has_s1w) count_slots([&](auto&& f) { for_seg_w(p, [&](uint32_t b, int, int, int) { f(b); }); }, p.n_slots_w);
      count_slots([&](auto&& f) { for_seg_u(p, [&](uint32_t b, int, int, int, int) { f(b); }); }, p.n_slots_u);
      count_slots([&](auto&& f) { for_seg_2(p, [&](uint32_t b, int, int, int, int, int) { f(b); }); }, p.n_slots_2);
      p.w_slots = kMaxWSlots;
      while (p.w_slots > 2 && tc_plan(p).total > kSmemCap) --p.w_slots;
      if (tc_plan(p).total <= kSmemCap && p.w_slots >= (sb == kSlotBytes ? 3 : 2)) break;
    }
    if (tc_plan(p).total > kSmemCap) { *why = "activation buffers of this layer do not fit shared memory next to a weight ring"; return false; }
  }
  return true;
}

bool tc_supported(const ModelDesc& md, const ForwardArgs& a, const char** why) {
  if (a.mask) { *why = "mask is an FP32-engine feature"; return false; }
  if ((a.h0 == nullptr) != (a.c0 == nullptr)) { *why = "initial_state needs both h and c"; return false; }
  if (a.flags & (SVDLSTM_GO_BACKWARDS | SVDLSTM_TIME_MAJOR)) { *why = "go_backwards / time_major are FP32-engine features"; return false; }
  if (!(a.flags & SVDLSTM_RETURN_SEQUENCES)) { *why = "return_sequences=False is an FP32-engine feature"; return false; }
  ModelDesc emd;
  if (!tc_effective_desc(md, emd, why)) return false;
  for (int l = 0; l < emd.n_layers; ++l) {
    TcLayerParams p;
    if (!tc_layer_params(emd, l, 32, p, why)) return false;
  }
  return true;
}

// one layer launch; 64-sequence tiles exist for units <= 256 only (32 cell states per unit block per thread)
template <int NUB, bool STREAM>
static int tc_launch_layer(const TcLayerParams& p, int n_cta, uint32_t smem_bytes, cudaStream_t stream) {
  if (p.ns == 64) {
    if constexpr (NUB <= 2) {
      SVD_CUDA_TRY(cudaFuncSetAttribute(lstm_tc_layer_kernel<NUB, STREAM, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
      lstm_tc_layer_kernel<NUB, STREAM, 64><<<n_cta, kTcThreads, smem_bytes, stream>>>(p);
      return 0;
    } else {
      set_error("tensor-core engine: 64-sequence tiles need units <= 256");
      return -1;
    }
  }
  if (p.h0 || p.h_n || p.c_n) {   // initial_state / return_state: the instantiation that carries the state plumbing
    SVD_CUDA_TRY(cudaFuncSetAttribute(lstm_tc_layer_kernel<NUB, STREAM, 32, kEpiWarps, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    lstm_tc_layer_kernel<NUB, STREAM, 32, kEpiWarps, true><<<n_cta, kTcThreads, smem_bytes, stream>>>(p);
    return 0;
  }
  SVD_CUDA_TRY(cudaFuncSetAttribute(lstm_tc_layer_kernel<NUB, STREAM, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  lstm_tc_layer_kernel<NUB, STREAM, 32><<<n_cta, kTcThreads, smem_bytes, stream>>>(p);
  return 0;
}

// units = 1024: 8 unit blocks, weights always streamed, 32-sequence tiles, 16 epilogue warps (cell state: 8 x 8 registers per thread)
static int tc_launch_layer_1024(const TcLayerParams& p, int n_cta, uint32_t smem_bytes, cudaStream_t stream) {
  SVD_REQUIRE(p.streaming && p.ns == 32, "tensor-core engine: units = 1024 needs a streamed weight ring and 32-sequence tiles");
  SVD_CUDA_TRY(cudaFuncSetAttribute(lstm_tc_layer_kernel<8, true, 32, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  lstm_tc_layer_kernel<8, true, 32, 16><<<n_cta, 32 * (4 + 16), smem_bytes, stream>>>(p);
  return 0;
}
static int tc_launch_pipe_1024(const void* pp, int n_cta, uint32_t smem_bytes, cudaStream_t stream);

// all layers in one cooperative launch (co-residency of every CTA is what makes the inter-layer waits safe)
template <int NUB, int NS, bool RAWX = false>
static int tc_launch_pipe(const TcPipeParams& pp, uint32_t smem_bytes, cudaStream_t stream) {
  SVD_CUDA_TRY(cudaFuncSetAttribute(lstm_tc_pipe_kernel<NUB, NS, kEpiWarps, RAWX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  void* args[] = {const_cast<TcPipeParams*>(&pp)};
  SVD_CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(lstm_tc_pipe_kernel<NUB, NS, kEpiWarps, RAWX>),
                                           dim3((unsigned)(pp.n_layers * pp.n_tiles)), dim3(kTcThreads), args, smem_bytes, stream));
  return 0;
}

static int tc_launch_pipe_1024(const void* pp_, int n_cta, uint32_t smem_bytes, cudaStream_t stream) {
  const TcPipeParams* pp = static_cast<const TcPipeParams*>(pp_);
  SVD_CUDA_TRY(cudaFuncSetAttribute(lstm_tc_pipe_kernel<8, 32, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  void* args[] = {const_cast<TcPipeParams*>(pp)};
  SVD_CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(lstm_tc_pipe_kernel<8, 32, 16>), dim3((unsigned)n_cta), dim3(32 * (4 + 16)), args,
                                           smem_bytes, stream));
  return 0;
}

static int run_tc_one(const ModelDesc& md, const ModelDesc& raw, TcState** state, bool weights_dirty, const ForwardArgs& a, cudaStream_t stream,
                      int* launches, int force_ns);

int run_tc(const ModelDesc& raw, TcState** state, bool weights_dirty, const ForwardArgs& a, cudaStream_t stream, int* launches) {
  ModelDesc emd;   // split layers as their merged equivalents (shapes; the tensors are built when the weights are packed)
  const char* why = "";
  SVD_REQUIRE(tc_effective_desc(raw, emd, &why), "tensor-core engine: %s", why);
  return run_tc_one(emd, raw, state, weights_dirty, a, stream, launches, 0);
}

// md = the effective (all-merged) description, raw = the handle's own
static int run_tc_one(const ModelDesc& md, const ModelDesc& raw, TcState** state, bool weights_dirty, const ForwardArgs& a, cudaStream_t stream,
                      int* launches, int force_ns) {
  const char* why = "";
  int nl = 0;
  if (*state == nullptr) {
    *state = new TcState();
    weights_dirty = true;
  }
  TcState* st = *state;
  const int L = md.n_layers;
  // Launch shape.  "pipe": all layers in ONE co-resident launch (CTA = layer x tile), possible when layers x tiles fits the
  // SMs; 64-sequence tiles make that true for twice the batch and halve the per-sequence cost of streamed weights.
  // "seq": one launch per layer, 32-sequence tiles, any batch.  SVDLSTM_TC_MODE=seq|pipe and SVDLSTM_TC_NS=32|64 override.
  int n_sm = 148;
  {
    int dev0 = 0;
    SVD_CUDA_TRY(cudaGetDevice(&dev0));
    SVD_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev0));
  }
  if (weights_dirty) st->ns = 0;   // the images follow the weights: rebuilt below
  const char* mode_env = getenv("SVDLSTM_TC_MODE");
  bool same_h = true;
  for (int l = 1; l < L; ++l) same_h = same_h && round_up(md.layers[l].units, 128) == round_up(md.layers[0].units, 128);
  bool ok64 = true;
  for (int l = 0; l < L; ++l) {
    TcLayerParams q;
    if (!tc_layer_params(md, l, 64, q, &why)) ok64 = false;
  }
  const char* ns_env = getenv("SVDLSTM_TC_NS");
  const bool allow_pipe = L >= 2 && same_h && !(mode_env && mode_env[0] == 's');
  int ns = 32;
  bool pipe = false;
  const bool has_state = a.h0 || a.h_n || a.c_n;
  if (has_state) {
    // initial_state / return_state: one launch per layer, 32-sequence tiles (the kernel instantiations that carry the state plumbing)
    SVD_REQUIRE(md.layers[0].units <= 512, "tensor-core engine: initial_state / return_state need units <= 512");
  } else if (ns_env || force_ns) {
    ns = ((force_ns ? force_ns : atoi(ns_env)) == 64 && ok64) ? 64 : 32;
    pipe = allow_pipe && L * ((a.B + ns - 1) / ns) <= n_sm;
  } else if (allow_pipe && L * ((a.B + 31) / 32) <= n_sm) {
    pipe = true;
  } else if (allow_pipe && ok64 && L * ((a.B + 63) / 64) <= n_sm) {
    pipe = true;
    ns = 64;
  }
  if (!pipe && allow_pipe && !ns_env && !force_ns && !has_state) {   // (state arrays are (layer, B, H): not sliceable by one pointer offset)
    // Batches too large for one co-resident launch (the rank x sequence sweep: 65 536 sequences): run them as chunks that ARE
    // pipelined -- layers x tiles <= SM count at the widest tile -- instead of layer-by-layer launches with 32-sequence tiles
    // (measured on C4: 82 ms -> see DESIGN.md section 5 per forward of 65 536 x 200).  Chunks are independent sequences.
    const int nsc = ok64 ? 64 : 32;
    const int chunk = (n_sm / L) * nsc;
    if (chunk >= nsc && a.B > chunk) {
      const int n_chunks = (a.B + chunk - 1) / chunk;
      const int per = round_up((a.B + n_chunks - 1) / n_chunks, nsc);
      int total = 0;
      for (int b0 = 0; b0 < a.B; b0 += per) {
        ForwardArgs sub = a;
        sub.B = imin(per, a.B - b0);
        sub.x = a.x + (size_t)b0 * a.T * md.input_dim;
        sub.y = a.y + (size_t)b0 * a.T * (md.n_out > 0 ? md.n_out : md.layers[L - 1].units);
        int nl_sub = 0;
        const int rc = run_tc_one(md, raw, state, weights_dirty && b0 == 0, sub, stream, &nl_sub, nsc);
        if (rc != 0) return rc;
        total += nl_sub;
      }
      *launches = total;
      return 0;
    }
  }
  if (st->ns != ns) weights_dirty = true;
  if (weights_dirty) {
    st->ns = ns;
    // effective factor blocks: 2-factor (ident) blocks are re-orthogonalised into 3-factor ones first (see tc_reorthogonalise)
    Block eff[kMaxLayers][2];
    std::vector<void*> temps;
    struct TempGuard {
      std::vector<void*>& t;
      cudaStream_t s;
      ~TempGuard() {
        if (!t.empty()) cudaStreamSynchronize(s);   // the pack kernels read them
        for (void* q : t) cudaFree(q);
      }
    } temp_guard{temps, stream};
    for (int l = 0; l < L; ++l)
      for (int j = 0; j < 2; ++j) {
        eff[l][j] = md.layers[l].blocks[j];
        if (raw.layers[l].n_blocks != 2) {
          if (int rc = tc_merge_side(raw.layers[l], j, stream, &eff[l][j], temps, &nl)) return rc;
        } else if (eff[l][j].ident) {
          const int rows = j == 0 ? md.layers[l].d_in : md.layers[l].units;
          if (int rc = tc_reorthogonalise(md.layers[l].blocks[j], rows, stream, &eff[l][j], temps, &nl)) return rc;
        }
      }
    for (int l = 0; l < L; ++l) {
      TcLayerImg& li = st->layers[l];
      TcLayerParams p;
      SVD_REQUIRE(tc_layer_params(md, l, ns, p, &why), "tensor-core engine: %s", why);
      if (li.wimg) cudaFree(li.wimg);
      if (li.bias) cudaFree(li.bias);
      li.wimg = nullptr;
      li.bias = nullptr;
      const uint32_t wbytes = p.segw_bytes + p.segu_bytes + p.seg2_bytes;
      SVD_CUDA_TRY(cudaMalloc(&li.wimg, wbytes));
      SVD_CUDA_TRY(cudaMalloc(&li.bias, sizeof(float) * 4 * p.H));
      // chunk table: the SAME iteration the kernel uses defines where every chunk lives
      std::vector<PackChunk> chunks;
      uint32_t off = 0;
      if (p.has_s1w)
        for_seg_w(p, [&](uint32_t b, int r0, int rows, int k0) {
          chunks.push_back(PackChunk{off, 0, (int16_t)rows, 64, (int16_t)r0, (int16_t)k0, 0, 0, 0});
          off += b;
        });
      for_seg_u(p, [&](uint32_t b, int, int r0, int rows, int k0) {
        chunks.push_back(PackChunk{off, 1, (int16_t)rows, 64, (int16_t)r0, (int16_t)k0, 0, 0, 0});
        off += b;
      });
      for_seg_2(p, [&](uint32_t b, int ub, int g, int part, int k0, int kc) {
        chunks.push_back(PackChunk{off, 2, 128, (int16_t)kc, 0, (int16_t)k0, (int16_t)ub, (int16_t)g, (int16_t)part});
        off += b;
      });
      if (li.chunks) {
        SVD_CUDA_TRY(cudaStreamSynchronize(stream));   // a previous forward may still be streaming from the old table
        cudaFree(li.chunks);
        li.chunks = nullptr;
      }
      SVD_CUDA_TRY(cudaMalloc(&li.chunks, sizeof(PackChunk) * chunks.size()));
      // pageable source: the copy is staged before the call returns, so `chunks` may die afterwards
      SVD_CUDA_TRY(cudaMemcpyAsync(li.chunks, chunks.data(), sizeof(PackChunk) * chunks.size(), cudaMemcpyHostToDevice, stream));
      p.chunks = li.chunks;
      if (li.slots) {
        cudaFree(li.slots);   // (the stream was synchronised above if a table existed)
        li.slots = nullptr;
      }
      p.slots = nullptr;
      if (p.streaming) {
        std::vector<uint32_t> slots;
        int seg = -1;
        uint32_t cur = 0, start = 0;
        auto flush = [&]() {
          if (cur) slots.push_back((start >> 8) | ((cur >> 8) << 16));
          cur = 0;
        };
        for (const PackChunk& c : chunks) {
          const uint32_t b = (uint32_t)c.rows * (uint32_t)c.kc * 2u;
          if (c.seg != seg || cur + b > p.slot_bytes) {
            flush();
            start = c.byte_off;
            seg = c.seg;
          }
          cur += b;
        }
        flush();
        SVD_REQUIRE((int)slots.size() == p.n_slots_w + p.n_slots_u + p.n_slots_2, "tensor-core engine: slot table mismatch (%d vs %d)",
                    (int)slots.size(), p.n_slots_w + p.n_slots_u + p.n_slots_2);
        SVD_CUDA_TRY(cudaMalloc(&li.slots, sizeof(uint32_t) * slots.size()));
        // (async like the chunk table: a blocking cudaMemcpy waits for everything queued before it -- in a rank sweep that is the
        //  previous model's whole forward, and the host then packs this model with the GPU idle)
        SVD_CUDA_TRY(cudaMemcpyAsync(li.slots, slots.data(), sizeof(uint32_t) * slots.size(), cudaMemcpyHostToDevice, stream));
        p.slots = li.slots;
      }
      SVD_REQUIRE((int)chunks.size() == p.n_chunks_w + p.n_chunks_u + p.n_chunks_2, "tensor-core engine: chunk table mismatch");
      const LayerDesc& Ld = md.layers[l];
      const Block bw_next = l + 1 < L ? eff[l + 1][0] : Block{};
      pack_wstream_kernel<<<(unsigned)chunks.size(), 256, 0, stream>>>(li.chunks, eff[l][0], eff[l][1], bw_next, p.x_r0, p.rx, p.H, p.Hu, Ld.d_in, p.ru_pad,
                                                                       md.dense_kernel, p.n_dense, md.n_out,
                                                                       reinterpret_cast<__half*>(li.wimg));
      pack_bias_kernel<<<(4 * p.H + 255) / 256, 256, 0, stream>>>(Ld.bias, p.H, p.Hu, li.bias);
      nl += 2;
      p.wimg = li.wimg;
      p.bias = li.bias;
      li.prm = p;
    }
    st->n_layers = L;
    SVD_CUDA_TRY(cudaGetLastError());
  }
  const int B = a.B, T = a.T;
  const int n_cta = (B + ns - 1) / ns;
  // workspaces (per device, shared by all handles)
  int dev = 0;
  SVD_CUDA_TRY(cudaGetDevice(&dev));
  SVD_REQUIRE(dev >= 0 && dev < 16, "tensor-core engine: device ordinal %d out of range", dev);
  TcWorkspace* ws = &g_tc_ws[dev];
  std::lock_guard<std::mutex> ws_lock(ws->mu);
  if (ws->done == nullptr) SVD_CUDA_TRY(cudaEventCreateWithFlags(&ws->done, cudaEventDisableTiming));
  else SVD_CUDA_TRY(cudaStreamWaitEvent(stream, ws->done, 0));   // the previous forward (any stream) owns the shared scratch until then
  const int Dpad = st->layers[0].prm.Kin;
  const size_t xbytes = (size_t)n_cta * T * act_tile_bytes(Dpad, ns);   // (allocated even when this forward reads x raw: a later one may not)
  if (ws->xseq_bytes < xbytes) {
    if (ws->xseq) {
      SVD_CUDA_TRY(cudaStreamSynchronize(stream));
      cudaFree(ws->xseq);
    }
    SVD_CUDA_TRY(cudaMalloc(&ws->xseq, xbytes));
    ws->xseq_bytes = xbytes;
  }
  // Hand-off rings (pipelined launch): the layers run a few steps apart (ring stages + one step of pipeline skew), so 32 steps per
  // tile are ample; C3 rank 128: 64 tiles x 32 x 16 KB = 32 MB, L2-resident (SVDLSTM_TC_RING=0: one slot per step, as in round 1).
  int ring = 32;
  if (const char* e = getenv("SVDLSTM_TC_RING")) ring = atoi(e);
  if (ring < 4 || ring >= T) ring = 0;
  for (int l = 0; l < L; ++l) {
    if (!st->layers[l].prm.store_h && !st->layers[l].prm.store_x) continue;
    const int steps = (pipe && l + 1 < L && ring > 0) ? ring : T;      // inter-layer hand-off of a pipelined launch: a ring
    const size_t hb = (size_t)n_cta * steps * (st->layers[l].prm.store_x ? act_tile_bytes(st->layers[l].prm.rx_pad, ns) : act_tile_bytes(st->layers[l].prm.H, ns));
    const int slot = pipe ? l : (l & 1);
    if (ws->seq_bytes[slot] < hb) {
      if (ws->seq[slot]) {
        SVD_CUDA_TRY(cudaStreamSynchronize(stream));
        cudaFree(ws->seq[slot]);
      }
      SVD_CUDA_TRY(cudaMalloc(&ws->seq[slot], hb));
      ws->seq_bytes[slot] = hb;
    }
  }
  if (pipe) {
    const size_t need = (size_t)2 * L * n_cta;     // [published | consumed]
    if (ws->progress_elems < need) {
      if (ws->progress) {
        SVD_CUDA_TRY(cudaStreamSynchronize(stream));
        cudaFree(ws->progress);
      }
      SVD_CUDA_TRY(cudaMalloc(&ws->progress, sizeof(int) * need));
      ws->progress_elems = need;
    }
    SVD_CUDA_TRY(cudaMemsetAsync(ws->progress, 0, sizeof(int) * need, stream));
  }
  // layer 0 reads the caller's float32 x itself (its input warp converts on the fly); SVDLSTM_TC_PACKX=1 restores the separate
  // packing pass + FP16 image of x (round 1: 2.8 % of the step, 368 MB of DRAM traffic per forward)
  // Only the 64-sequence pipelined launch of units <= 256 with a long weight stream (issue-/port-bound, ranks >= ~96) has a RAWX
  // build: there the pass is pure overhead; the epilogue-bound ranks keep the bulk-copy loader (see tc_layer_body).
  bool raw_x = false;
  if (pipe && ns == 64 && getenv("SVDLSTM_TC_PACKX") == nullptr) {
    const TcLayerParams& lp = st->layers[L - 1].prm;
    // (a forward that follows its own upload -- a.x_ready -- takes the raw-x build at every rank: a few per cent of kernel time
    //  against the whole upload in series)
    raw_x = lp.streaming && ((lp.n_chunks_u + lp.n_chunks_2) >= 32 || a.x_ready != nullptr) && st->layers[0].prm.streaming;
  }
  if (!raw_x && a.x_ready != nullptr) {
    set_error("tensor-core engine: this model / batch does not take the raw-x pipelined launch (needed for a streamed input upload)");
    return -3;
  }
  if (!raw_x) {
    const int D = md.input_dim;
    const int TT = D <= 16 ? 8 : (D <= 32 ? 4 : 2);
    pack_x_kernel<<<dim3((T + TT - 1) / TT, n_cta), 256, sizeof(float) * ns * TT * D, stream>>>(a.x, B, T, D, Dpad, TT, ns, ws->xseq);
    ++nl;
  }
  static long long* dbg_buf = nullptr;
  const char* dbg_env = getenv("SVDLSTM_TC_TIMELINE");
  if (dbg_env && !dbg_buf) {
    SVD_CUDA_TRY(cudaMalloc(&dbg_buf, sizeof(long long) * kDbgPerLayer * kMaxLayers));
    SVD_CUDA_TRY(cudaMemset(dbg_buf, 0, sizeof(long long) * kDbgPerLayer * kMaxLayers));
  }
  TcPipeParams pp{};
  uint32_t pipe_smem = 0;
  for (int l = 0; l < L; ++l) {
    TcLayerParams p = st->layers[l].prm;
    p.T = T;
    p.B = B;
    p.dbg = dbg_env ? dbg_buf + (size_t)l * kDbgPerLayer : nullptr;
    const int in_slot = pipe ? l - 1 : ((l - 1) & 1), out_slot = pipe ? l : (l & 1);
    p.in_seq = (l == 0) ? ws->xseq : ws->seq[in_slot];
    p.in_ring = (pipe && l > 0) ? ring : 0;
    p.out_ring = (pipe && l + 1 < L) ? ring : 0;
    p.x_raw = (l == 0 && raw_x) ? a.x : nullptr;
    p.x_ready = (l == 0 && raw_x) ? a.x_ready : nullptr;
    p.x_dim = md.input_dim;
    p.out_seq = (p.store_h || p.store_x) ? ws->seq[out_slot] : nullptr;
    p.y = a.y;
    p.dense_bias = md.dense_bias;
    {
      size_t soff = 0;   // state arrays: [layer][B][units]
      for (int i = 0; i < l; ++i) soff += (size_t)B * md.layers[i].units;
      p.h0 = a.h0 ? a.h0 + soff : nullptr;
      p.c0 = a.c0 ? a.c0 + soff : nullptr;
      p.h_n = a.h_n ? a.h_n + soff : nullptr;
      p.c_n = a.c_n ? a.c_n + soff : nullptr;
    }
    const TcSmemPlan sp = tc_plan(p);
    if (pipe) {
      pp.layer[l] = p;
      pipe_smem = sp.total > pipe_smem ? sp.total : pipe_smem;
      continue;
    }
    int lrc = -1;
    switch ((p.H / 128) * 2 + (p.streaming ? 1 : 0)) {
      case 2: lrc = tc_launch_layer<1, false>(p, n_cta, sp.total, stream); break;
      case 3: lrc = tc_launch_layer<1, true>(p, n_cta, sp.total, stream); break;
      case 4: lrc = tc_launch_layer<2, false>(p, n_cta, sp.total, stream); break;
      case 5: lrc = tc_launch_layer<2, true>(p, n_cta, sp.total, stream); break;
      case 6: lrc = tc_launch_layer<3, false>(p, n_cta, sp.total, stream); break;
      case 7: lrc = tc_launch_layer<3, true>(p, n_cta, sp.total, stream); break;
      case 8: lrc = tc_launch_layer<4, false>(p, n_cta, sp.total, stream); break;
      case 9: lrc = tc_launch_layer<4, true>(p, n_cta, sp.total, stream); break;
      case 17: lrc = tc_launch_layer_1024(p, n_cta, sp.total, stream); break;
      default: set_error("tensor-core engine: unsupported units %d", p.H);
    }
    if (lrc != 0) return lrc;
    ++nl;
  }
  if (pipe) {
    pp.n_layers = L;
    pp.n_tiles = n_cta;
    pp.progress = ws->progress;
    pp.consumed = ws->progress + (size_t)L * n_cta;
    int lrc = -1;
    switch ((st->layers[0].prm.H / 128) * 2 + (ns == 64 ? 1 : 0)) {
      case 2: lrc = tc_launch_pipe<1, 32>(pp, pipe_smem, stream); break;
      case 3: lrc = raw_x ? tc_launch_pipe<1, 64, true>(pp, pipe_smem, stream) : tc_launch_pipe<1, 64>(pp, pipe_smem, stream); break;
      case 4: lrc = tc_launch_pipe<2, 32>(pp, pipe_smem, stream); break;
      case 5: lrc = raw_x ? tc_launch_pipe<2, 64, true>(pp, pipe_smem, stream) : tc_launch_pipe<2, 64>(pp, pipe_smem, stream); break;   // (16 epilogue warps measured slower here: 5.75 vs 5.55 ms at rank 64)
      case 6: lrc = tc_launch_pipe<3, 32>(pp, pipe_smem, stream); break;
      case 8: lrc = tc_launch_pipe<4, 32>(pp, pipe_smem, stream); break;
      case 16: lrc = tc_launch_pipe_1024(&pp, L * n_cta, pipe_smem, stream); break;
      default: set_error("tensor-core engine: unsupported units %d for the pipelined launch", md.layers[0].units);
    }
    if (lrc != 0) return lrc;
    ++nl;
  }
  SVD_REQUIRE(!raw_x || (pipe && ns == 64 && st->layers[0].prm.H <= 256), "tensor-core engine: internal error (raw-x forward without a raw-x kernel)");
  if (st->layers[L - 1].prm.store_h) {
    const uint8_t* last_seq = ws->seq[pipe ? L - 1 : ((L - 1) & 1)];
    if (md.n_out > 0)      // Dense top that did not fit the S1u tiles
      dense_top_kernel<<<dim3((unsigned)T, (unsigned)n_cta), 256, 0, stream>>>(last_seq, B, T, st->layers[L - 1].prm.H, st->layers[L - 1].prm.Hu, ns,
                                                                              md.dense_kernel, md.dense_bias, md.n_out, a.y);
    else                   // no Dense top: the output is the last hidden sequence itself
      unpack_out_kernel<<<dim3(128, n_cta), 256, 0, stream>>>(last_seq, B, T, st->layers[L - 1].prm.H, st->layers[L - 1].prm.Hu, ns, a.y);
    ++nl;
  }
  SVD_CUDA_TRY(cudaGetLastError());
  SVD_CUDA_TRY(cudaEventRecord(ws->done, stream));
  if (dbg_env) {   // debugging aid: dump the per-step timeline of CTA 0 (cycles relative to the step's first stamp)
    static long long host[kDbgPerLayer * kMaxLayers];
    SVD_CUDA_TRY(cudaStreamSynchronize(stream));
    SVD_CUDA_TRY(cudaMemcpy(host, dbg_buf, sizeof(long long) * kDbgPerLayer * L, cudaMemcpyDeviceToHost));
    for (int l = 0; l < L; ++l)
      for (int t = 20; t < 24 && t < T; ++t) {
        const long long* r = host + (size_t)l * kDbgPerLayer + (size_t)t * 16;
        fprintf(stderr, "[tc timeline] layer %d step %d (stream=%d slots=%d):", l, t, st->layers[l].prm.streaming, st->layers[l].prm.w_slots);
        for (int i = 0; i < 15; ++i) fprintf(stderr, " %lld", r[i] ? r[i] - r[2] : -1);
        fprintf(stderr, "  | step period %lld\n", r[2] - (r - 16)[2]);
      }
    for (int l = 0; l < L; ++l) {   // issue time of every weight chunk of step 20, relative to that step's S1 commit
      const long long* r = host + (size_t)l * kDbgPerLayer;
      fprintf(stderr, "[tc chunks] layer %d step 20:", l);
      for (int i = 0; i < 250 && r[64 * 16 + i]; ++i) fprintf(stderr, " %lld", r[64 * 16 + i] - r[20 * 16 + 2]);
      fprintf(stderr, "\n");
    }
  }
  *launches = nl;
  return 0;
}

}  // namespace svdlstm
