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

struct ForwardArgs {
  const float* x;
  float* y;
  const float* h0;
  const float* c0;
  float* h_n;
  float* c_n;
  const uint8_t* mask;
  int B, T, flags;
};

// engines (k1_*.cu); each returns 0 / error code and the number of kernels it launched
int run_general(const ModelDesc& host_md, const ModelDesc* dev_md, const ForwardArgs& a,
                cudaStream_t stream, int* launches);
bool wavefront_supported(const ModelDesc& md, const ForwardArgs& a);
int run_wavefront(const ModelDesc& host_md, const ModelDesc* dev_md, const ForwardArgs& a,
                  cudaStream_t stream, int* launches);
bool tc_supported(const ModelDesc& md, const ForwardArgs& a, const char** why);
struct TcState;  // packed bf16 weights owned by the handle
int run_tc(const ModelDesc& host_md, TcState** state, bool weights_dirty, const ForwardArgs& a,
                cudaStream_t stream, int* launches);
void tc_free(TcState* s);

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace svdlstm
