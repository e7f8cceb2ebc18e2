/*
 * svdlstm.h -- C-ABI of the B200-native SVD-factored LSTM hot path (libsvdlstm.so).
 *
 * The reference (dncoble/LSTM-acceleration-with-singular-value-decomposition) is pure
 * Python on Keras; it has no FFI.  Its drop-in boundary is the Keras layer protocol of
 * code/svd_classes_v3.py.  The Python classes of this repo keep that protocol (same names,
 * constructor kwargs, get_weights() orderings) and bind the entry points below through
 * ctypes.  Each entry point cites the reference interface it replaces (file:line relative
 * to the reference root).
 *
 * Conventions
 *   - plain pointers and sizes only; every `const float*` / `float*` is a DEVICE pointer owned by
 *     the caller (PyTorch tensors in this repo); row-major, float32, row-vector convention
 *     (svd_classes_v3.py:128), gate order i,f,c,o along the 4H axis (:144).
 *   - weight pointers are BORROWED until the next set_*_weights for that layer or destroy.
 *   - all work is enqueued on the caller's stream (`stream` is a cudaStream_t passed as void*);
 *     no hidden synchronisation unless stated.
 *   - return 0 = ok; <0 = argument / shape / state error; >0 = cudaError_t.  The message is in
 *     svdlstm_last_error() (thread-local).
 *   - there is no CPU path: without a CUDA device every compute entry point fails.
 */
#ifndef SVDLSTM_H_
#define SVDLSTM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svdlstm_model_s* svdlstm_handle;

/* forward flags (Keras LSTM kwargs used by SingularLSTM.call, svd_classes_v3.py:385-437) */
#define SVDLSTM_RETURN_SEQUENCES      1   /* :428-431 */
#define SVDLSTM_GO_BACKWARDS          2   /* :413     */
#define SVDLSTM_TIME_MAJOR            4   /* :417     */
#define SVDLSTM_ZERO_OUTPUT_FOR_MASK  8   /* :418     */

/* engines */
#define SVDLSTM_ENGINE_AUTO     0   /* regime switch (north_star): the tensor-core engine when the call is dense enough for it
                                       -- batch >= SVDLSTM_TC_MIN_BATCH, widest layer >= SVDLSTM_TC_MIN_UNITS, and the model /
                                       call is one it supports -- else the FP32 choice below.  The environment variable
                                       SVDLSTM_STRICT_FP32=1 turns AUTO into ENGINE_FP32 process-wide.                          */
#define SVDLSTM_ENGINE_GENERAL  1   /* FP32 CUDA-core batched persistent kernel (any shape)                                    */
#define SVDLSTM_ENGINE_WAVEFRONT 2  /* FP32 register-resident warp-per-layer wavefront (H,D,r<=32)                             */
#define SVDLSTM_ENGINE_TC       3   /* tcgen05 tensor-core persistent kernel: FP16 operands, FP32 accumulate + cell state
                                       (reduced precision; 3-/2-factor or full cells, merged or split -- a split layer's gate
                                       blocks are merged when the weights are packed --, units <= 512 or <= 1024 padded to
                                       128-row tiles, ranks <= 256, return_sequences, no mask)                                 */
#define SVDLSTM_ENGINE_FP32     4   /* strict FP32 (the 1e-5 parity path): wavefront when the model fits it, else general     */
#define SVDLSTM_TC_MIN_BATCH    128
#define SVDLSTM_TC_MIN_UNITS    64

/* ---- model handle ------------------------------------------------------------------------
 * A handle describes a stack of L LSTM layers (+ optional Dense top), i.e. what the
 * reference builds as keras.Sequential([SingularLSTM...]+[Dense]) (svd_classes_v3.py:471-540,
 * 554-598, 605-676).  Host metadata only; no device work.                                   */
int svdlstm_create(svdlstm_handle* out, int n_layers, int input_dim, const int* units);
void svdlstm_destroy(svdlstm_handle h);

/* Stock LSTMCell weights  get_weights() = [W (D,4H), U (H,4H), b (4H,)]
 * (what make_LSTM_singular_model reads at svd_classes_v3.py:557).                          */
int svdlstm_set_full_weights(svdlstm_handle h, int layer, const float* W, const float* U,
                             const float* b);

/* SingularLSTMCell weights in get_weights() order (svd_classes_v3.py:113):
 *   w[0]=kernel sigma_w, w[1]=recurrent_kernel sigma_u, w[2]=w_left, w[3]=w_right,
 *   w[4]=u_left, w[5]=u_right, w[6]=bias.
 * merged (:35-56,81-112): sigma_w (1,k_w) w_left (D,k_w) w_right (k_w,4H); same for u.
 * split  (:58-107): sigma_w (1,4k_w) w_left (D,4k_w) w_right (k_w,4H) (gate g = axis-1 quarter g).
 * k_w/k_u are the kept ranks (per gate when split).                                          */
int svdlstm_set_singular_weights(svdlstm_handle h, int layer, int merged,
                                 const float* const* w, int k_w, int k_u);

/* ReducedLSTMCell weights in get_weights() order (svd_classes_v3.py:278, :308-315):
 *   merged: w = [w_left (D,r_w), w_right (r_w,4H-r_w), u_left (H,r_u), u_right (r_u,4H-r_u), bias],
 *           ranks = [r_w, r_u]
 *   split : w = [w_left_g (D,rw_g), w_right_g (rw_g,H-rw_g), u_left_g (H,ru_g), u_right_g (ru_g,H-ru_g)]
 *           for g in i,f,c,o, then bias;  ranks = [rw_i,ru_i, rw_f,ru_f, rw_c,ru_c, rw_o,ru_o]   */
int svdlstm_set_reduced_weights(svdlstm_handle h, int layer, int merged,
                                const float* const* w, const int* ranks);

/* Dense / TimeDistributed(Dense) top: kernel (H_last, n_out), bias (n_out)
 * (svd_classes_v3.py:532-539, 590-597, 670-675).  kernel==NULL removes the top.            */
int svdlstm_set_dense_top(svdlstm_handle h, const float* kernel, const float* bias, int n_out);

/* SingularLSTM.call / Sequential.predict (svd_classes_v3.py:385-437; svd_acceleration_v3.py:151):
 * one persistent launch runs all T steps of all layers (+ Dense top).
 *   x     (B,T,D)  [or (T,B,D) with TIME_MAJOR]
 *   h0,c0 NULL (zero state, :393) or the per-layer states concatenated: layer l occupies
 *         B*H_l floats after those of layers < l.  h_n,c_n same layout, may be NULL.
 *   mask  NULL or (B,T) bytes (non-zero = valid step), Keras mask semantics.
 *   y     (B,T,n) with RETURN_SEQUENCES else (B,n); n = n_out if a Dense top is set, else H_last.
 *         With GO_BACKWARDS outputs are in processing order (Keras does not re-reverse).     */
int svdlstm_forward(svdlstm_handle h, const float* x, int B, int T, float* y,
                    const float* h0, const float* c0, float* h_n, float* c_n,
                    const uint8_t* mask, int flags, int engine, void* stream);

/* Pinned (page-locked) host buffers for the inputs / outputs of predict(); write_combined != 0: cudaHostAllocWriteCombined
 * (DMA reads are not snooped by the CPU caches -- faster host->device streams when several GPUs share a socket; slow for the
 * CPU to read back, so for input staging only).                                                                           */
int svdlstm_host_alloc(void** out, size_t bytes, int write_combined);
int svdlstm_host_free(void* p);

/* Number of kernels the last forward on this handle launched, and which engine ran. */
/* One forward from a PINNED host array with the upload INSIDE the forward (latency of a single `rmodel.predict(X)`,
 * svd_acceleration_v3.py:150-151): x_host (B,T,D) is copied to x_dev in n_slices equal time slices (0: slices of 64, 64, 128, 256 ...
 * steps -- a short first one, long fast rows afterwards) on `copy_stream`; the tensor-core
 * kernel is launched on `stream` as soon as the first slice has landed and follows the upload through a progress word (its layer-0
 * input warp reads x(t) straight from x_dev).  Same result as svdlstm_forward(ENGINE_TC) on the uploaded array.  Returns -3 if the
 * tensor-core engine does not take this model at all (nothing was enqueued: upload and run it the plain way), -4 if the launch this
 * batch takes cannot follow an upload (low ranks, small batches): the slices ARE enqueued, so the caller waits for `copy_stream`
 * and calls svdlstm_forward on x_dev.  x_host must stay unmodified until `copy_stream` has drained.                              */
int svdlstm_forward_streamed_input(svdlstm_handle h, const float* x_host, float* x_dev, int B, int T, float* y, int n_slices,
                                   void* copy_stream, void* stream);
int svdlstm_last_launches(svdlstm_handle h);
int svdlstm_last_engine(svdlstm_handle h);

/* Weight-count helper == sum(a.size for a in model.get_weights()) without the Dense top
 * (svd_acceleration_v3.py:160-166).                                                          */
int64_t svdlstm_count_weights(svdlstm_handle h);

/* ---- batched one-sided Jacobi SVD -----------------------------------------------------------
 * Replaces np.linalg.svd(mat, full_matrices=False) at svd_classes_v3.py:491,562 and
 * old_versions/svd_classes.py:10,15,231.  A: batch x (m,n) row-major contiguous.
 * Outputs (k=min(m,n)): U batch x (m,k), S batch x (k) descending, Vt batch x (k,n).
 * Any of U/Vt may be NULL (values only).  sweeps (device int[batch], may be NULL) receives
 * the number of Jacobi sweeps used, NEGATED when the matrix hit the sweep cap (40) without
 * converging (its factors are then not orthogonal to working precision).  Internally float64. */
int svdlstm_svd_jacobi_batched(const float* A, int batch, int m, int n, float* U, float* S,
                               float* Vt, int* sweeps, void* stream);

/* B = (U_r * S_r) V1 ; C = V1^-1 V2  with V_r = [V1 (r,r) | V2 (r,n-r)]
 * (svd_classes_v3.py:622-626, 656-660).  Inputs already sliced to the kept rank r:
 * U_r (m,r) with row stride ldu, S_r (r), V_r (r,n) with row stride ldv.  Outputs B (m,r), C (r,n-r)
 * contiguous.  pivot_ratio (device float[1], may be NULL) = min|pivot| / max|pivot| of the
 * partial-pivoting elimination of V1 (small => V1 ill-conditioned; the reference calls inv()
 * unguarded).                                                                                 */
int svdlstm_reduce_factors(const float* U_r, int ldu, const float* S_r, const float* V_r, int ldv,
                           int m, int r, int n, float* B, float* C, float* pivot_ratio,
                           void* stream);

/* The same for a batch of matrices (every factor of a model, or of a whole rank sweep) in ONE launch.  */
typedef struct {
  const float* U_r;
  const float* S_r;
  const float* V_r;
  float* B;
  float* C;
  float* pivot_ratio;
  int32_t ldu, ldv, m, r, n;
} svdlstm_reduce_item;
int svdlstm_reduce_factors_batched(const svdlstm_reduce_item* items, int n_items, void* stream);

/* out (m,n; row stride ldo) = (A (m,k; lda) * diag(scale (k) or NULL)) . B (k,n; ldb) + bias (n) or NULL.
 * The dense products of the path outside the recurrent kernels: the rank-truncated reconstruction
 * A_r = (U * s_r) V with k = the kept rank (old_versions/svd_classes.py:9-12, 210-217) and a stand-alone
 * Dense / TimeDistributed(Dense) layer (svd_classes_v3.py:532-539).                                     */
int svdlstm_scaled_matmul(const float* A, int lda, const float* scale, const float* B, int ldb, const float* bias,
                          int m, int k, int n, float* out, int ldo, void* stream);

/* ---- fused Hoyer + orthogonality penalties ----------------------------------------------------
 * One launch evaluates every regulariser of a model: HoyerRegularizer.__call__
 * (svd_classes_v3.py:460-462) on sigma vectors and keras OrthogonalRegularizer(mode='rows')
 * (call sites :514,:573) on the factor matrices.  items[i] is an (rows,cols) row-major matrix
 * (a vector is rows=1).  out[4*i..4*i+3] (device doubles) =
 *   { sum|x|, sum x^2, sum_{i!=j}|Pn_ij| (Gram of L2-normalised rows), ||X X^T - I||_F^2 }.
 * gram[i]==0 skips the two Gram sums (sigma vectors).  columns[i]!=0 uses columns instead of rows. */
typedef struct {
  const float* data;
  int32_t rows, cols, ld;
  int32_t gram;      /* 0: L1/L2 sums only; 1: also Gram sums */
  int32_t columns;   /* 0: rows mode (Keras default here); 1: columns mode */
} svdlstm_penalty_item;
int svdlstm_penalties(const svdlstm_penalty_item* items, int n_items, double* out, void* stream);

/* ---- rank-sweep squared error -----------------------------------------------------------------
 * sse[r] = sum_i (pred[r,i]-target[i])^2, deterministic two-stage reduction in float64 -- the
 * device half of the RMSE of svd_acceleration_v3.py:187-190 / old_versions/svd_acceleration.py:79-81.
 * pred (n_ranks, n) contiguous, target (n), sse device double[n_ranks].                         */
int svdlstm_sweep_sse(const float* pred, const float* target, int n_ranks, int64_t n, double* sse,
                      void* stream);

/* ---- training step (Hoyer fine-tune) -------------------------------------------------------------
 * Replaces smodel.compile(loss="mse", optimizer="adam") + smodel.fit(...) of the reference driver
 * (svd_acceleration_v3.py:111-128) for models of SingularLSTMCells + Dense top.  Trainable tensors as in the
 * reference: sigma_w / sigma_u always (svd_classes_v3.py:40,47), the factor matrices and the bias only where
 * train_uv[layer] != 0 (:49-56,:102-112), the Dense top always.  The flat parameter / gradient vector is laid
 * out per layer as [sigma_w | sigma_u | w_left | w_right | u_left | u_right | bias] (get_weights() order, :113),
 * then [dense kernel | dense bias]; svdlstm_trainer_layout returns the 7 L + 3 start offsets.
 *   gradients     forward with cache + back-propagation through time of the mean-squared error of one mini-batch
 *                 (x (B,T,D); y_true (B,T,n) with return_sequences else (B,n); device pointers).  grad_out NULL =
 *                 loss only (validation).  Deterministic: per-sequence partial gradients, summed in batch order.
 *   regularizers  adds d penalty / d w and the penalty itself for HoyerRegularizer (kind 1, :455-462) and
 *                 OrthogonalRegularizer(mode='rows') (kind 2, call sites :514,:573); tensor_index = 7 l + {0..5}.
 *   adam          one Keras-Adam update IN PLACE in the caller's weight tensors.                                */
typedef struct svdlstm_trainer_s* svdlstm_trainer;
int svdlstm_trainer_create(svdlstm_handle h, const int* train_uv, svdlstm_trainer* out);
void svdlstm_trainer_destroy(svdlstm_trainer t);
int64_t svdlstm_trainer_num_params(svdlstm_trainer t);
int svdlstm_trainer_layout(svdlstm_trainer t, int64_t* offsets);
int svdlstm_trainer_gradients(svdlstm_trainer t, const float* x, const float* y_true, int B, int T, int return_sequences,
                              float* grad_out, float* loss_out, void* stream);
int svdlstm_trainer_regularizers(svdlstm_trainer t, const int* tensor_index, const int* kind, const float* coef, int n,
                                 float* grad, float* loss, void* stream);
int svdlstm_trainer_adam(svdlstm_trainer t, const float* grad, float lr, float beta1, float beta2, float eps, void* stream);

/* ---- real-time batch-1 stream -------------------------------------------------------------------
 * The reference's deployment setting: a stateful LSTM (svd_classes_v3.py:421-426) fed one frame every
 * 400-500 us (train_full_model_v4.py:14-16), i.e. `model.predict` called per sample
 * (svd_acceleration_v3.py:151).  svdlstm_stream_open launches ONE persistent kernel for the handle's
 * model (it must fit the wavefront engine: units, input_dim, ranks <= 32; input_dim <= 30; <= 15
 * outputs); afterwards a sample costs no CUDA call at all: svdlstm_stream_step writes the input_dim
 * floats of x_t (HOST pointer) into a ring in host-mapped pinned memory, the kernel polls it, runs all
 * layers + the Dense top from registers and writes y_t back into host memory, where the call picks it
 * up (blocking; y_t is a HOST pointer to n_out -- or units_last -- floats).  State (h, c) persists
 * between calls.  The kernel parks its state and leaves after idle_ms (<= 0: 50 ms) without a sample
 * -- a forgotten stream can neither pin an SM nor stall a cudaDeviceSynchronize for long -- and the
 * next step relaunches it transparently; the same happens after the handle's weights are re-bound.
 * One caller thread per stream.
 *   svdlstm_stream_run   feeds n samples paced at period_us (0 = back to back) from native code and
 *                        records the per-sample latency (write of x_t -> y_t visible) in latency_us
 *                        (may be NULL): the measurement loop of bench.py, free of interpreter overhead.
 *   svdlstm_stream_reset sets the state (NULL, NULL = zeros; else [layer][units] arrays, host or device).
 *   svdlstm_stream_state reads the state back the same way.                                          */
typedef struct svdlstm_stream_s* svdlstm_stream;
int svdlstm_stream_open(svdlstm_handle h, int idle_ms, svdlstm_stream* out);
int svdlstm_stream_step(svdlstm_stream s, const float* x_t, float* y_t);
int svdlstm_stream_run(svdlstm_stream s, const float* x, int n, double period_us, float* y, float* latency_us);
int svdlstm_stream_reset(svdlstm_stream s, const float* h0, const float* c0);
int svdlstm_stream_state(svdlstm_stream s, float* h_n, float* c_n);
int svdlstm_stream_launches(svdlstm_stream s);
int svdlstm_stream_close(svdlstm_stream s);

const char* svdlstm_last_error(void);
const char* svdlstm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SVDLSTM_H_ */
