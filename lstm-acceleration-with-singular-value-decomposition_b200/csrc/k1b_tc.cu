// K1b: tcgen05 BF16 tensor-core persistent recurrent kernel (placeholder until the kernel lands).
#include "common.cuh"

namespace svdlstm {

struct TcState {
  int unused;
};

bool tc_supported(const ModelDesc& md, const ForwardArgs& a, const char** why) {
  (void)md;
  (void)a;
  *why = "tensor-core engine not built into this library yet";
  return false;
}

int run_tc_bf16(const ModelDesc& md, TcState** state, bool weights_dirty, const ForwardArgs& a, cudaStream_t stream, int* launches) {
  (void)md; (void)state; (void)weights_dirty; (void)a; (void)stream; (void)launches;
  set_error("tensor-core engine not built into this library yet");
  return -3;
}

void tc_free(TcState* s) { delete s; }

}  // namespace svdlstm
