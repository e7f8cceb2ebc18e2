// Shared internals of libsvdlstm.so: error plumbing + the device-side model description.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/svdlstm.h"

namespace svdlstm {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define SVD_CUDA_TRY(expr)                                         \
  do {                                                             \
    cudaError_t _e = (expr);                                       \
    if (_e != cudaSuccess) return ::svdlstm::cuda_fail(_e, #expr); \
  } while (0)

#define SVD_REQUIRE(cond, ...)           \
  do {                                   \
    if (!(cond)) {                       \
      ::svdlstm::set_error(__VA_ARGS__); \
      return -1;                         \
    }                                    \
  } while (0)

constexpr int kMaxLayers = 8;
constexpr int kMaxBlocks = 8;  // split form: 4 gates x {W,U}

// One factor block of one weight matrix of one layer.  Every cell form of the reference
// reduces to a list of these (DESIGN.md "canonical block form"):
//   p[k]  = scale[k] * sum_i in[i] * left[i*left_ld + k]          (stage 1; left==nullptr: p = in)
//   ident==0: z[out0 + c]        += sum_k p[k] * right[k*right_ld + c],  c < ncols   (3-factor / full)
//   ident==1: z[out0 + k]        += p[k],  k < rank                                   (2-factor:
//             z[out0 + rank + c] += sum_k p[k] * right[k*right_ld + c],  c < ncols     [a | a C])
struct Block {
  const float* left;
  const float* scale;
  const float* right;
  int left_ld, right_ld;
  int rank;
  int ncols;
  int out0;
  int ident;
  int from_h;  // 0: input is the layer input x_t / h_{l-1}(t); 1: input is h_l(t-1)
  int p_off;   // offset of this block's p vector inside the layer's p scratch
};

struct LayerDesc {
  int d_in, units;
  int n_blocks;
  int p_total;
  const float* bias;
  Block blocks[kMaxBlocks];
};

struct ModelDesc {
  int n_layers;
  int input_dim;
  int n_out;  // 0: no dense top
  const float* dense_kernel;
  const float* dense_bias;
  LayerDesc layers[kMaxLayers];
};

struct TcState;  // packed FP16 weight-stream images owned by the handle (k1b_tc.cu)

struct ForwardArgs {
  const float* x;
  float* y;
  const float* h0;
  const float* c0;
  float* h_n;
  float* c_n;
  const uint8_t* mask;
  int B, T, flags;
  const int* x_ready;   // optional (tensor-core raw-x launch only): device word = number of leading time steps of x that have landed
                        // (x is still being uploaded, time slice by time slice, while the kernel runs); nullptr: x is complete
};

// ---- real-time stream (k1_wavefront.cu <STREAM>, stream.cu): rings in host-mapped pinned memory -------------------------------
// Sample s travels in slot s % kStreamSlots tagged s+1.  A tag is only trusted together with the data of its OWN 64-byte line
// (the PCIe root complex reads / writes a cache line atomically; x86 stores become visible in program order), so the input
// slot carries the tag in both of its lines.
constexpr int kStreamSlots = 64;
struct StreamSlotIn { uint32_t w[32]; };    // line A = [tag, x0..x14], line B = [tag, x15..x29]
struct StreamSlotOut { uint32_t w[16]; };   // [tag, y0..y14]
struct StreamCtl {
  volatile uint32_t stop;          // host -> device: leave at the next poll
  uint32_t pad0[15];
  volatile uint32_t exited;        // device -> host: generation of the launch that has left (written last)
  volatile uint32_t exit_reason;   // 1 = stop, 2 = idle
  volatile uint32_t consumed;      // samples consumed by all launches so far
  uint32_t pad1[13];
};
struct StreamArgs {
  StreamSlotIn* in;
  StreamSlotOut* out;
  StreamCtl* ctl;
  uint32_t first_seq;              // samples consumed before this launch (state was parked after them)
  uint32_t generation;
  unsigned long long idle_ns;      // leave after this long without a sample
};

}  // namespace svdlstm

// the opaque handle of include/svdlstm.h
struct svdlstm_model_s {
  svdlstm::ModelDesc md;             // host copy (device pointers inside)
  bool layer_set[svdlstm::kMaxLayers];
  svdlstm::ModelDesc* dev_md;        // device copy, re-uploaded when dirty
  svdlstm::ModelDesc* pinned_md;     // pinned staging so the upload is stream-ordered and async
  cudaEvent_t md_event;              // recorded after every forward: the last reader of dev_md / pinned_md (whatever its stream)
  unsigned long long md_version;     // bumped by every upload (a running real-time stream notices re-bound weights)
  bool dirty;
  bool tc_dirty;
  struct svdlstm::TcState* tc;
  int last_launches;
  int last_engine;
  int64_t n_weights[svdlstm::kMaxLayers];
  int* xr_dev;                       // svdlstm_forward_streamed_input: progress word on the device ...
  int* xr_host;                      // ... and the pinned table of the values the copy stream publishes (slice ends)
  cudaEvent_t xr_event;              // first slice landed
  cudaEvent_t xr_done;               // whole upload enqueued by the last call has drained (the table may be rewritten)
};

namespace svdlstm {

int upload_model_desc(svdlstm_model_s* h, cudaStream_t stream);

// engines (k1_*.cu); each returns 0 / error code and the number of kernels it launched
int run_general(const ModelDesc& host_md, const ModelDesc* dev_md, const ForwardArgs& a,
                cudaStream_t stream, int* launches);
bool wavefront_supported(const ModelDesc& md, const ForwardArgs& a);
int run_wavefront(const ModelDesc& host_md, const ModelDesc* dev_md, const ForwardArgs& a,
                  cudaStream_t stream, int* launches);
bool stream_supported(const ModelDesc& md, const char** why);
int launch_wavefront_stream(const ModelDesc& md, const ModelDesc* dev_md, float* state_h, float* state_c, const StreamArgs& sa,
                            cudaStream_t stream);
bool tc_supported(const ModelDesc& md, const ForwardArgs& a, const char** why);
int run_tc(const ModelDesc& host_md, TcState** state, bool weights_dirty, const ForwardArgs& a,
                cudaStream_t stream, int* launches);
void tc_free(TcState* s);

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace svdlstm
