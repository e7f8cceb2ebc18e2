// Real-time batch-1 service: host half of the persistent-kernel protocol (device half: k1_wavefront.cu <STREAM = true>).
//
// The reference's deployment target is one 16-sample frame every 400-500 us (train_full_model_v4.py:14-16) through a stateful
// LSTM (svd_classes_v3.py:421-426).  A launch + H2D + D2H + synchronize per sample costs >100 us through any CUDA API; here the
// kernel is launched ONCE and then fed through rings in host-mapped pinned memory: per sample the host writes 128 bytes, the
// device polls them over PCIe, runs the L layers + Dense top out of registers, and writes the tagged prediction back into
// host memory, where the caller spins on it.  No CUDA call is made per sample.
//
// The kernel parks its state and leaves on `stop` or after `idle_ms` without a sample (so that a forgotten stream can neither
// pin an SM nor block a cudaDeviceSynchronize for ever); the next step relaunches it transparently from the parked state.
#include <sched.h>
#include <string.h>
#include <time.h>

#include <new>

#include "common.cuh"

using namespace svdlstm;

struct svdlstm_stream_s {
  svdlstm_model_s* h;
  cudaStream_t cstream;          // private non-blocking stream the persistent kernel lives on
  StreamSlotIn* in;              // host-mapped rings (host pointers; dev_* are the device aliases)
  StreamSlotOut* out;
  StreamCtl* ctl;
  StreamSlotIn* dev_in;
  StreamSlotOut* dev_out;
  StreamCtl* dev_ctl;
  float* state_h;                // parked (h, c) of every layer, [layer][units], device memory
  float* state_c;
  int state_floats;
  int D, n_y;
  uint32_t submitted, completed; // samples written to / read back from the rings
  uint32_t generation;           // launches so far
  bool running;                  // a launch of generation `generation` may still be resident
  unsigned long long md_version; // model description the resident kernel staged
  unsigned long long idle_ns;
  int relaunches;
};

namespace {

inline double now_us() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#endif
}

int stream_launch(svdlstm_stream_s* s) {
  svdlstm_model_s* h = s->h;
  if (int e = upload_model_desc(h, s->cstream)) return e;
  s->md_version = h->md_version;
  s->ctl->stop = 0;
  ++s->generation;
  StreamArgs sa{s->dev_in, s->dev_out, s->dev_ctl, s->completed, s->generation, s->idle_ns};
  __atomic_thread_fence(__ATOMIC_SEQ_CST);
  if (int e = launch_wavefront_stream(h->md, h->dev_md, s->state_h, s->state_c, sa, s->cstream)) return e;
  if (!h->md_event) SVD_CUDA_TRY(cudaEventCreateWithFlags(&h->md_event, cudaEventDisableTiming));
  SVD_CUDA_TRY(cudaEventRecord(h->md_event, s->cstream));   // dev_md is read by the kernel's prologue
  s->running = true;
  ++s->relaunches;
  return 0;
}

// Ask the resident kernel to leave and wait until it has (its state is parked in device memory afterwards).
int stream_quiesce(svdlstm_stream_s* s) {
  if (!s->running) return 0;
  s->ctl->stop = 1;
  __atomic_thread_fence(__ATOMIC_SEQ_CST);
  SVD_CUDA_TRY(cudaStreamSynchronize(s->cstream));
  s->running = false;
  return 0;
}

inline void write_sample(svdlstm_stream_s* s, const float* x) {
  StreamSlotIn* slot = &s->in[s->submitted % kStreamSlots];
  const uint32_t tag = (s->submitted & 0x7fffffffu) + 1u;
  uint32_t buf[32];
  memset(buf, 0, sizeof(buf));
  for (int i = 0; i < s->D; ++i) {
    uint32_t u;
    memcpy(&u, &x[i], 4);
    buf[i < 15 ? 1 + i : 2 + i] = u;      // line A holds x0..x14 in words 1..15, line B x15..x29 in words 17..31
  }
  volatile uint32_t* w = slot->w;
  for (int i = 1; i < 16; ++i) w[i] = buf[i];
  for (int i = 17; i < 32; ++i) w[i] = buf[i];
  __atomic_thread_fence(__ATOMIC_RELEASE);   // data before tags (x86: compiler barrier; stores are not reordered)
  w[0] = tag;
  w[16] = tag;
  ++s->submitted;
}

// Wait for the prediction of sample `completed`; relaunches the kernel if it left (idle) before consuming the sample.
int read_result(svdlstm_stream_s* s, float* y, double timeout_us) {
  const uint32_t want = (s->completed & 0x7fffffffu) + 1u;
  volatile uint32_t* w = s->out[s->completed % kStreamSlots].w;
  const double t0 = now_us();
  unsigned spins = 0;
  while (true) {
    if (w[0] == want) break;
    if ((++spins & 0xFF) == 0) {
      if (s->ctl->exited == s->generation && w[0] != want) {
        // the kernel left (idle time-out, or an error) without this sample: resume from the parked state
        cudaError_t e = cudaStreamSynchronize(s->cstream);
        if (e != cudaSuccess) return cuda_fail(e, "real-time stream kernel");
        s->running = false;
        if (w[0] == want) break;
        if (int rc = stream_launch(s)) return rc;
      }
      if (now_us() - t0 > timeout_us) {
        set_error("svdlstm_stream: no prediction for sample %u within %.0f ms", s->completed, timeout_us * 1e-3);
        return -4;
      }
    }
    cpu_relax();
  }
  __atomic_thread_fence(__ATOMIC_ACQUIRE);
  for (int i = 0; i < s->n_y; ++i) {
    const uint32_t u = w[1 + i];
    memcpy(&y[i], &u, 4);
  }
  ++s->completed;
  return 0;
}

int ensure_running(svdlstm_stream_s* s) {
  if (s->h->dirty || s->md_version != s->h->md_version) {   // weights were re-bound: the kernel holds stale copies in registers
    if (int e = stream_quiesce(s)) return e;
  }
  if (s->running && s->ctl->exited == s->generation) {
    SVD_CUDA_TRY(cudaStreamSynchronize(s->cstream));
    s->running = false;
  }
  if (!s->running) return stream_launch(s);
  return 0;
}

}  // namespace

extern "C" {

int svdlstm_stream_open(svdlstm_handle h, int idle_ms, svdlstm_stream* out) {
  SVD_REQUIRE(h != nullptr && out != nullptr, "svdlstm_stream_open: null argument");
  for (int l = 0; l < h->md.n_layers; ++l) SVD_REQUIRE(h->layer_set[l], "svdlstm_stream_open: weights of layer %d were never set", l);
  const char* why = "";
  if (!stream_supported(h->md, &why)) {
    set_error("svdlstm_stream_open: %s", why);
    return -3;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("svdlstm_stream_open: no CUDA device (this library has no CPU path)");
    return -2;
  }
  svdlstm_stream_s* s = new (std::nothrow) svdlstm_stream_s();
  SVD_REQUIRE(s != nullptr, "svdlstm_stream_open: out of host memory");
  memset(s, 0, sizeof(*s));
  s->h = h;
  s->D = h->md.input_dim;
  s->n_y = h->md.n_out > 0 ? h->md.n_out : h->md.layers[h->md.n_layers - 1].units;
  s->idle_ns = (unsigned long long)(idle_ms > 0 ? idle_ms : 50) * 1000000ull;
  for (int l = 0; l < h->md.n_layers; ++l) s->state_floats += h->md.layers[l].units;
  cudaError_t e = cudaSuccess;
  void* ring = nullptr;
  const size_t ring_bytes = sizeof(StreamSlotIn) * kStreamSlots + sizeof(StreamSlotOut) * kStreamSlots + sizeof(StreamCtl);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->cstream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaHostAlloc(&ring, ring_bytes, cudaHostAllocMapped | cudaHostAllocPortable);
  if (e == cudaSuccess) {
    memset(ring, 0, ring_bytes);
    s->in = reinterpret_cast<StreamSlotIn*>(ring);
    s->out = reinterpret_cast<StreamSlotOut*>(s->in + kStreamSlots);
    s->ctl = reinterpret_cast<StreamCtl*>(s->out + kStreamSlots);
    void* dring = nullptr;
    e = cudaHostGetDevicePointer(&dring, ring, 0);
    s->dev_in = reinterpret_cast<StreamSlotIn*>(dring);
    s->dev_out = reinterpret_cast<StreamSlotOut*>(s->dev_in + kStreamSlots);
    s->dev_ctl = reinterpret_cast<StreamCtl*>(s->dev_out + kStreamSlots);
  }
  if (e == cudaSuccess) e = cudaMalloc(&s->state_h, sizeof(float) * 2 * s->state_floats);
  if (e == cudaSuccess) {
    s->state_c = s->state_h + s->state_floats;
    e = cudaMemset(s->state_h, 0, sizeof(float) * 2 * s->state_floats);   // zero initial state (svd_classes_v3.py:393)
  }
  if (e != cudaSuccess) {
    if (s->state_h) cudaFree(s->state_h);
    if (ring) cudaFreeHost(ring);
    if (s->cstream) cudaStreamDestroy(s->cstream);
    delete s;
    return cuda_fail(e, "svdlstm_stream_open");
  }
  *out = s;
  return 0;
}

int svdlstm_stream_step(svdlstm_stream s, const float* x_t, float* y_t) {
  SVD_REQUIRE(s != nullptr && x_t != nullptr && y_t != nullptr, "svdlstm_stream_step: null argument");
  if (int e = ensure_running(s)) return e;
  write_sample(s, x_t);
  return read_result(s, y_t, 5e6);
}

int svdlstm_stream_run(svdlstm_stream s, const float* x, int n, double period_us, float* y, float* latency_us) {
  SVD_REQUIRE(s != nullptr && x != nullptr && y != nullptr && n >= 0, "svdlstm_stream_run: bad argument");
  if (int e = ensure_running(s)) return e;
  // best effort: a real-time producer runs at a real-time priority (a descheduled feeder thread shows up as a millisecond outlier
  // that has nothing to do with the device); silently skipped without CAP_SYS_NICE
  const int old_policy = sched_getscheduler(0);
  sched_param old_param{};
  sched_getparam(0, &old_param);
  sched_param rt_param{};
  rt_param.sched_priority = 10;
  const bool rt = period_us > 0 && sched_setscheduler(0, SCHED_FIFO, &rt_param) == 0;
  struct Restore {
    bool on;
    int pol;
    sched_param par;
    ~Restore() {
      if (on) sched_setscheduler(0, pol, &par);
    }
  } restore{rt, old_policy, old_param};
  const double t_start = now_us();
  for (int i = 0; i < n; ++i) {
    if (period_us > 0) {
      const double due = t_start + (double)i * period_us;
      while (now_us() < due) cpu_relax();      // a paced producer: sample i becomes available at its deadline
    }
    const double t_in = now_us();
    write_sample(s, x + (size_t)i * s->D);
    if (int e = read_result(s, y + (size_t)i * s->n_y, 5e6)) return e;
    if (latency_us) latency_us[i] = (float)(now_us() - t_in);
  }
  return 0;
}

int svdlstm_stream_reset(svdlstm_stream s, const float* h0, const float* c0) {
  SVD_REQUIRE(s != nullptr, "svdlstm_stream_reset: null stream");
  SVD_REQUIRE((h0 == nullptr) == (c0 == nullptr), "svdlstm_stream_reset: h0 and c0 must both be given or both be NULL");
  if (int e = stream_quiesce(s)) return e;
  if (h0) {
    SVD_CUDA_TRY(cudaMemcpy(s->state_h, h0, sizeof(float) * s->state_floats, cudaMemcpyDefault));
    SVD_CUDA_TRY(cudaMemcpy(s->state_c, c0, sizeof(float) * s->state_floats, cudaMemcpyDefault));
  } else {
    SVD_CUDA_TRY(cudaMemset(s->state_h, 0, sizeof(float) * 2 * s->state_floats));
  }
  return 0;
}

int svdlstm_stream_state(svdlstm_stream s, float* h_n, float* c_n) {
  SVD_REQUIRE(s != nullptr && h_n != nullptr && c_n != nullptr, "svdlstm_stream_state: null argument");
  if (int e = stream_quiesce(s)) return e;
  SVD_CUDA_TRY(cudaMemcpy(h_n, s->state_h, sizeof(float) * s->state_floats, cudaMemcpyDefault));
  SVD_CUDA_TRY(cudaMemcpy(c_n, s->state_c, sizeof(float) * s->state_floats, cudaMemcpyDefault));
  return 0;
}

int svdlstm_stream_launches(svdlstm_stream s) { return s ? s->relaunches : -1; }

int svdlstm_stream_close(svdlstm_stream s) {
  if (!s) return 0;
  int rc = stream_quiesce(s);
  cudaFree(s->state_h);
  cudaFreeHost(s->in);
  cudaStreamDestroy(s->cstream);
  delete s;
  return rc;
}

}  // extern "C"
