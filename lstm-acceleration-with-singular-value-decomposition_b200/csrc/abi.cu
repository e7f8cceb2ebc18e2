// extern "C" surface of libsvdlstm.so: model handle, weight binding, forward dispatch.
// See include/svdlstm.h for the contract and the reference interfaces each entry replaces.
#include <stdlib.h>
#include <string.h>

#include <new>

#include "common.cuh"

namespace svdlstm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return (int)e;
}

}  // namespace svdlstm

using namespace svdlstm;

namespace svdlstm {
// Device copy of the model description, re-uploaded (stream-ordered, through a pinned staging copy) when weights were re-bound.
int upload_model_desc(svdlstm_model_s* h, cudaStream_t stream) {
  if (!h->dirty) return 0;
  if (!h->dev_md) SVD_CUDA_TRY(cudaMalloc(&h->dev_md, sizeof(ModelDesc)));
  if (!h->pinned_md) SVD_CUDA_TRY(cudaMallocHost(&h->pinned_md, sizeof(ModelDesc)));
  // An earlier forward -- possibly on ANOTHER stream -- may still be reading dev_md, or its upload may still be reading the
  // pinned staging copy: wait for the event recorded after that forward, not for the current stream.
  if (h->md_event) SVD_CUDA_TRY(cudaEventSynchronize(h->md_event));
  memcpy(h->pinned_md, &h->md, sizeof(ModelDesc));
  SVD_CUDA_TRY(cudaMemcpyAsync(h->dev_md, h->pinned_md, sizeof(ModelDesc), cudaMemcpyHostToDevice, stream));
  h->dirty = false;
  ++h->md_version;
  return 0;
}
}  // namespace svdlstm

extern "C" {

const char* svdlstm_last_error(void) { return g_err; }
const char* svdlstm_version(void) { return "svdlstm-b200 0.1 (sm_100a)"; }

int svdlstm_create(svdlstm_handle* out, int n_layers, int input_dim, const int* units) {
  SVD_REQUIRE(out != nullptr && units != nullptr, "svdlstm_create: null argument");
  SVD_REQUIRE(n_layers >= 1 && n_layers <= kMaxLayers, "svdlstm_create: n_layers=%d not in [1,%d]", n_layers, kMaxLayers);
  SVD_REQUIRE(input_dim >= 1, "svdlstm_create: input_dim=%d", input_dim);
  svdlstm_model_s* m = new (std::nothrow) svdlstm_model_s();
  SVD_REQUIRE(m != nullptr, "svdlstm_create: out of host memory");
  memset(&m->md, 0, sizeof(ModelDesc));
  m->md.n_layers = n_layers;
  m->md.input_dim = input_dim;
  int d = input_dim;
  for (int l = 0; l < n_layers; ++l) {
    if (units[l] < 1) {
      delete m;
      set_error("svdlstm_create: units[%d]=%d", l, units[l]);
      return -1;
    }
    m->md.layers[l].d_in = d;
    m->md.layers[l].units = units[l];
    m->layer_set[l] = false;
    m->n_weights[l] = 0;
    d = units[l];
  }
  m->dev_md = nullptr;
  m->xr_dev = nullptr;
  m->xr_host = nullptr;
  m->xr_event = nullptr;
  m->xr_done = nullptr;
  m->pinned_md = nullptr;
  m->md_event = nullptr;
  m->md_version = 0;
  m->dirty = true;
  m->tc_dirty = true;
  m->tc = nullptr;
  m->last_launches = 0;
  m->last_engine = 0;
  *out = m;
  return 0;
}

void svdlstm_destroy(svdlstm_handle h) {
  if (!h) return;
  if (h->dev_md) cudaFree(h->dev_md);
  if (h->pinned_md) cudaFreeHost(h->pinned_md);
  if (h->md_event) cudaEventDestroy(h->md_event);
  if (h->tc) tc_free(h->tc);
  if (h->xr_dev) cudaFree(h->xr_dev);
  if (h->xr_host) cudaFreeHost(h->xr_host);
  if (h->xr_event) cudaEventDestroy(h->xr_event);
  if (h->xr_done) cudaEventDestroy(h->xr_done);
  delete h;
}

static int check_layer(svdlstm_handle h, int layer, const char* fn) {
  SVD_REQUIRE(h != nullptr, "%s: null handle", fn);
  SVD_REQUIRE(layer >= 0 && layer < h->md.n_layers, "%s: layer %d out of range [0,%d)", fn, layer, h->md.n_layers);
  return 0;
}

static void finish_layer(svdlstm_handle h, int layer) {
  LayerDesc& L = h->md.layers[layer];
  int p = 0;
  for (int b = 0; b < L.n_blocks; ++b) {
    L.blocks[b].p_off = p;
    p += L.blocks[b].rank;
  }
  L.p_total = p;
  h->layer_set[layer] = true;
  h->dirty = true;
  h->tc_dirty = true;
}

int svdlstm_set_full_weights(svdlstm_handle h, int layer, const float* W, const float* U, const float* b) {
  if (int e = check_layer(h, layer, "svdlstm_set_full_weights")) return e;
  SVD_REQUIRE(W && U && b, "svdlstm_set_full_weights: null weight pointer");
  LayerDesc& L = h->md.layers[layer];
  const int H = L.units, D = L.d_in;
  L.n_blocks = 2;
  L.bias = b;
  Block& bw = L.blocks[0];
  bw = Block{nullptr, nullptr, W, 0, 4 * H, D, 4 * H, 0, 0, 0, 0};
  Block& bu = L.blocks[1];
  bu = Block{nullptr, nullptr, U, 0, 4 * H, H, 4 * H, 0, 0, 1, 0};
  h->n_weights[layer] = (int64_t)D * 4 * H + (int64_t)H * 4 * H + 4 * H;
  finish_layer(h, layer);
  return 0;
}

int svdlstm_set_singular_weights(svdlstm_handle h, int layer, int merged, const float* const* w, int k_w, int k_u) {
  if (int e = check_layer(h, layer, "svdlstm_set_singular_weights")) return e;
  SVD_REQUIRE(w != nullptr, "svdlstm_set_singular_weights: null weight list");
  for (int i = 0; i < 7; ++i) SVD_REQUIRE(w[i] != nullptr, "svdlstm_set_singular_weights: weight %d is null", i);
  LayerDesc& L = h->md.layers[layer];
  const int H = L.units, D = L.d_in;
  SVD_REQUIRE(k_w >= 1 && k_u >= 1, "svdlstm_set_singular_weights: ranks must be >= 1 (got %d,%d)", k_w, k_u);
  const float *s_w = w[0], *s_u = w[1], *w_l = w[2], *w_r = w[3], *u_l = w[4], *u_r = w[5];
  L.bias = w[6];
  if (merged) {
    SVD_REQUIRE(k_w <= (D < 4 * H ? D : 4 * H) && k_u <= H, "svdlstm_set_singular_weights: merged ranks (%d,%d) exceed min(D,4H)=%d / H=%d", k_w, k_u, (D < 4 * H ? D : 4 * H), H);
    L.n_blocks = 2;
    L.blocks[0] = Block{w_l, s_w, w_r, k_w, 4 * H, k_w, 4 * H, 0, 0, 0, 0};
    L.blocks[1] = Block{u_l, s_u, u_r, k_u, 4 * H, k_u, 4 * H, 0, 0, 1, 0};
    h->n_weights[layer] = (int64_t)k_w * (1 + D + 4 * H) + (int64_t)k_u * (1 + H + 4 * H) + 4 * H;
  } else {
    SVD_REQUIRE(k_w <= (D < H ? D : H) && k_u <= H, "svdlstm_set_singular_weights: split ranks (%d,%d) exceed min(D,H)=%d / H=%d", k_w, k_u, (D < H ? D : H), H);
    L.n_blocks = 8;
    for (int g = 0; g < 4; ++g) {
      // gate g = axis-1 quarter g of every concatenated array (svd_classes_v3.py:165-170,207-212)
      L.blocks[g] = Block{w_l + g * k_w, s_w + g * k_w, w_r + g * H, 4 * k_w, 4 * H, k_w, H, g * H, 0, 0, 0};
      L.blocks[4 + g] = Block{u_l + g * k_u, s_u + g * k_u, u_r + g * H, 4 * k_u, 4 * H, k_u, H, g * H, 0, 1, 0};
    }
    h->n_weights[layer] = (int64_t)4 * k_w * (1 + D + H) + (int64_t)4 * k_u * (1 + H + H) + 4 * H;
  }
  finish_layer(h, layer);
  return 0;
}

int svdlstm_set_reduced_weights(svdlstm_handle h, int layer, int merged, const float* const* w, const int* ranks) {
  if (int e = check_layer(h, layer, "svdlstm_set_reduced_weights")) return e;
  SVD_REQUIRE(w != nullptr && ranks != nullptr, "svdlstm_set_reduced_weights: null argument");
  LayerDesc& L = h->md.layers[layer];
  const int H = L.units, D = L.d_in;
  int64_t cnt = 4 * H;
  if (merged) {
    const int rw = ranks[0], ru = ranks[1];
    SVD_REQUIRE(rw >= 1 && rw <= 4 * H && ru >= 1 && ru <= 4 * H, "svdlstm_set_reduced_weights: merged ranks (%d,%d) not in [1,4H=%d]", rw, ru, 4 * H);
    SVD_REQUIRE(w[0] && w[2] && w[4], "svdlstm_set_reduced_weights: null weight pointer");
    SVD_REQUIRE((w[1] || rw == 4 * H) && (w[3] || ru == 4 * H), "svdlstm_set_reduced_weights: null right factor");
    L.n_blocks = 2;
    L.blocks[0] = Block{w[0], nullptr, w[1], rw, 4 * H - rw, rw, 4 * H - rw, 0, 1, 0, 0};
    L.blocks[1] = Block{w[2], nullptr, w[3], ru, 4 * H - ru, ru, 4 * H - ru, 0, 1, 1, 0};
    L.bias = w[4];
    cnt += (int64_t)rw * (D + 4 * H - rw) + (int64_t)ru * (H + 4 * H - ru);
  } else {
    L.n_blocks = 8;
    for (int g = 0; g < 4; ++g) {
      const int rw = ranks[2 * g], ru = ranks[2 * g + 1];
      SVD_REQUIRE(rw >= 1 && rw <= H && ru >= 1 && ru <= H, "svdlstm_set_reduced_weights: split ranks (%d,%d) of gate %d not in [1,H=%d]", rw, ru, g, H);
      const float *wl = w[4 * g], *wr = w[4 * g + 1], *ul = w[4 * g + 2], *ur = w[4 * g + 3];
      SVD_REQUIRE(wl && ul && (wr || rw == H) && (ur || ru == H), "svdlstm_set_reduced_weights: null weight pointer (gate %d)", g);
      L.blocks[g] = Block{wl, nullptr, wr, rw, H - rw, rw, H - rw, g * H, 1, 0, 0};
      L.blocks[4 + g] = Block{ul, nullptr, ur, ru, H - ru, ru, H - ru, g * H, 1, 1, 0};
      cnt += (int64_t)rw * (D + H - rw) + (int64_t)ru * (H + H - ru);
    }
    SVD_REQUIRE(w[16] != nullptr, "svdlstm_set_reduced_weights: null bias");
    L.bias = w[16];
  }
  h->n_weights[layer] = cnt;
  finish_layer(h, layer);
  return 0;
}

int svdlstm_set_dense_top(svdlstm_handle h, const float* kernel, const float* bias, int n_out) {
  SVD_REQUIRE(h != nullptr, "svdlstm_set_dense_top: null handle");
  if (kernel == nullptr) {
    h->md.n_out = 0;
    h->md.dense_kernel = nullptr;
    h->md.dense_bias = nullptr;
  } else {
    SVD_REQUIRE(n_out >= 1 && n_out <= 64, "svdlstm_set_dense_top: n_out=%d not in [1,64]", n_out);
    h->md.n_out = n_out;
    h->md.dense_kernel = kernel;
    h->md.dense_bias = bias;
  }
  h->dirty = true;
  h->tc_dirty = true;
  return 0;
}

/* Pinned host staging buffers for predict(): write_combined != 0 asks for cudaHostAllocWriteCombined memory -- not snooped
 * by the CPU caches during DMA, which is what limits host->device throughput when several GPUs stream their inputs from one
 * socket (slow for the CPU to READ, so only for buffers the host fills and the GPU consumes). */
int svdlstm_host_alloc(void** out, size_t bytes, int write_combined) {
  SVD_REQUIRE(out != nullptr && bytes > 0, "svdlstm_host_alloc: bad argument");
  SVD_CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)));
  return 0;
}
int svdlstm_host_free(void* p) {
  if (p) SVD_CUDA_TRY(cudaFreeHost(p));
  return 0;
}

int64_t svdlstm_count_weights(svdlstm_handle h) {
  if (!h) return -1;
  int64_t s = 0;
  for (int l = 0; l < h->md.n_layers; ++l) s += h->n_weights[l];
  return s;
}

int svdlstm_last_launches(svdlstm_handle h) { return h ? h->last_launches : -1; }
int svdlstm_last_engine(svdlstm_handle h) { return h ? h->last_engine : -1; }

int svdlstm_forward(svdlstm_handle h, const float* x, int B, int T, float* y, const float* h0, const float* c0,
                    float* h_n, float* c_n, const uint8_t* mask, int flags, int engine, void* stream_) {
  SVD_REQUIRE(h != nullptr, "svdlstm_forward: null handle");
  SVD_REQUIRE(x != nullptr && y != nullptr, "svdlstm_forward: null x / y");
  SVD_REQUIRE(B >= 1 && T >= 1, "svdlstm_forward: B=%d T=%d must be >= 1", B, T);
  SVD_REQUIRE((h0 == nullptr) == (c0 == nullptr), "svdlstm_forward: h0 and c0 must both be given or both be NULL");
  for (int l = 0; l < h->md.n_layers; ++l) SVD_REQUIRE(h->layer_set[l], "svdlstm_forward: weights of layer %d were never set", l);
  int ndev = 0;
  cudaError_t de = cudaGetDeviceCount(&ndev);
  if (de != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("svdlstm_forward: no CUDA device (this library has no CPU path)");
    return -2;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  ForwardArgs a{x, y, h0, c0, h_n, c_n, mask, B, T, flags};
  int launches = 0;
  int rc = 0;
  int used = engine;
  if (engine == SVDLSTM_ENGINE_AUTO) {
    // Regime switch: the thin contractions only fill 128-row x 32-sequence tensor-core tiles when batch x 4H is dense enough;
    // below that the FP32 latency engines win and keep full precision.
    static const bool strict = [] { const char* e = getenv("SVDLSTM_STRICT_FP32"); return e && e[0] && e[0] != '0'; }();
    int max_units = 0;
    for (int l = 0; l < h->md.n_layers; ++l) max_units = h->md.layers[l].units > max_units ? h->md.layers[l].units : max_units;
    const char* why = "";
    if (!strict && B >= SVDLSTM_TC_MIN_BATCH && max_units >= SVDLSTM_TC_MIN_UNITS && tc_supported(h->md, a, &why)) used = SVDLSTM_ENGINE_TC;
    else used = SVDLSTM_ENGINE_FP32;
  }
  if (used == SVDLSTM_ENGINE_FP32) used = wavefront_supported(h->md, a) ? SVDLSTM_ENGINE_WAVEFRONT : SVDLSTM_ENGINE_GENERAL;
  switch (used) {
    case SVDLSTM_ENGINE_GENERAL:
      // (the device copy of the descriptor -- a device + a pinned allocation per handle -- only for the engines that read it: a
      //  rank sweep builds hundreds of handles that only ever run on the tensor-core engine, which packs its own images)
      if (int e = upload_model_desc(h, stream)) return e;
      rc = run_general(h->md, h->dev_md, a, stream, &launches);
      break;
    case SVDLSTM_ENGINE_WAVEFRONT:
      SVD_REQUIRE(wavefront_supported(h->md, a), "svdlstm_forward: wavefront engine needs units,input_dim,ranks <= 32, <= %d layers, n_out <= 1, no mask, and factors within the register budget", 6);
      if (int e = upload_model_desc(h, stream)) return e;
      rc = run_wavefront(h->md, h->dev_md, a, stream, &launches);
      break;
    case SVDLSTM_ENGINE_TC: {
      const char* why = "";
      if (!tc_supported(h->md, a, &why)) {
        set_error("svdlstm_forward: tensor-core engine unsupported for this model/call: %s", why);
        return -3;
      }
      rc = run_tc(h->md, &h->tc, h->tc_dirty, a, stream, &launches);
      if (rc == 0) h->tc_dirty = false;
      break;
    }
    default:
      set_error("svdlstm_forward: unknown engine %d", engine);
      return -1;
  }
  if (rc == 0) {
    if (!h->md_event) SVD_CUDA_TRY(cudaEventCreateWithFlags(&h->md_event, cudaEventDisableTiming));
    SVD_CUDA_TRY(cudaEventRecord(h->md_event, stream));
  }
  h->last_launches = launches;
  h->last_engine = used;
  return rc;
}

constexpr int kMaxInputSlices = 64;

int svdlstm_forward_streamed_input(svdlstm_handle h, const float* x_host, float* x_dev, int B, int T, float* y, int n_slices, void* copy_stream_,
                                   void* stream_) {
  SVD_REQUIRE(h != nullptr, "svdlstm_forward_streamed_input: null handle");
  SVD_REQUIRE(x_host != nullptr && x_dev != nullptr && y != nullptr, "svdlstm_forward_streamed_input: null x_host / x_dev / y");
  SVD_REQUIRE(B >= 1 && T >= 1, "svdlstm_forward_streamed_input: B=%d T=%d must be >= 1", B, T);
  SVD_REQUIRE(n_slices >= 0 && n_slices <= kMaxInputSlices, "svdlstm_forward_streamed_input: n_slices=%d not in [0,%d]", n_slices, kMaxInputSlices);
  SVD_REQUIRE(copy_stream_ != stream_, "svdlstm_forward_streamed_input: the upload needs a stream of its own");
  for (int l = 0; l < h->md.n_layers; ++l) SVD_REQUIRE(h->layer_set[l], "svdlstm_forward_streamed_input: weights of layer %d were never set", l);
  cudaStream_t cs = (cudaStream_t)copy_stream_, stream = (cudaStream_t)stream_;
  ForwardArgs a{x_dev, y, nullptr, nullptr, nullptr, nullptr, nullptr, B, T, SVDLSTM_RETURN_SEQUENCES, nullptr};
  const char* why = "";
  if (!tc_supported(h->md, a, &why)) {
    set_error("svdlstm_forward_streamed_input: tensor-core engine unsupported for this model/call: %s", why);
    return -3;
  }
  if (!h->xr_dev) SVD_CUDA_TRY(cudaMalloc(&h->xr_dev, sizeof(int)));
  if (!h->xr_host) {
    SVD_CUDA_TRY(cudaMallocHost(&h->xr_host, sizeof(int) * (kMaxInputSlices + 1)));
  }
  if (!h->xr_event) SVD_CUDA_TRY(cudaEventCreateWithFlags(&h->xr_event, cudaEventDisableTiming));
  if (!h->xr_done) SVD_CUDA_TRY(cudaEventCreateWithFlags(&h->xr_done, cudaEventDisableTiming));
  else SVD_CUDA_TRY(cudaEventSynchronize(h->xr_done));   // the previous call's publishes still read the pinned table until then
  // the previous forward of this handle (any stream) may still be polling the progress word / reading the table's DMA source
  if (h->md_event) SVD_CUDA_TRY(cudaStreamWaitEvent(cs, h->md_event, 0));
  const int D = h->md.input_dim;
  // n_slices > 0: equal slices (an even number of steps each).  n_slices == 0: 64, 64, 128, 256, ... steps -- the kernel starts behind
  // a short first slice and the later ones are long rows, which the copy engine moves faster (a 2-D copy of 8 KB rows reaches
  // ~35 GB/s, a contiguous one ~55).
  const int per = n_slices > 0 ? ((T + n_slices - 1) / n_slices + 1) & ~1 : 0;
  h->xr_host[0] = 0;
  SVD_CUDA_TRY(cudaMemcpyAsync(h->xr_dev, h->xr_host, sizeof(int), cudaMemcpyHostToDevice, cs));
  const size_t pitch = sizeof(float) * (size_t)T * D;
  int k = 0;
  for (int t0 = 0, len = 64; t0 < T; ++k) {
    const int step = per > 0 ? per : len;
    if (per == 0 && k >= 1) len *= 2;
    const int t1 = t0 + step < T ? t0 + step : T;
    SVD_CUDA_TRY(cudaMemcpy2DAsync(x_dev + (size_t)t0 * D, pitch, x_host + (size_t)t0 * D, pitch, sizeof(float) * (size_t)(t1 - t0) * D, (size_t)B,
                                   cudaMemcpyHostToDevice, cs));
    h->xr_host[k + 1] = t1;
    SVD_CUDA_TRY(cudaMemcpyAsync(h->xr_dev, h->xr_host + k + 1, sizeof(int), cudaMemcpyHostToDevice, cs));
    if (k == 0) SVD_CUDA_TRY(cudaEventRecord(h->xr_event, cs));
    t0 = t1;
  }
  SVD_CUDA_TRY(cudaEventRecord(h->xr_done, cs));
  SVD_CUDA_TRY(cudaStreamWaitEvent(stream, h->xr_event, 0));
  a.x_ready = h->xr_dev;
  int launches = 0;
  int rc = run_tc(h->md, &h->tc, h->tc_dirty, a, stream, &launches);
  if (rc == 0) {
    h->tc_dirty = false;
    if (!h->md_event) SVD_CUDA_TRY(cudaEventCreateWithFlags(&h->md_event, cudaEventDisableTiming));
    SVD_CUDA_TRY(cudaEventRecord(h->md_event, stream));
    h->last_launches = launches;
    h->last_engine = SVDLSTM_ENGINE_TC;
  } else if (rc == -3) {
    rc = -4;   // the launch this batch takes cannot follow an upload -- but the slices ARE on their way (unlike -3 above)
  }
  return rc;
}

}  // extern "C"
