"""Sequential container + the reference's model builders ("rank-reduction API").

    make_LSTM_singular_model        code/svd_classes_v3.py:548-598
    make_split_LSTM_singular_model  code/svd_classes_v3.py:469-540
    make_LSTM_reduced_model         code/svd_classes_v3.py:604-676   (+ keyword-only ``rank=``)

The SVDs run on device (K2, batched one-sided Jacobi) instead of np.linalg.svd (:491,:562); the
B / C construction runs on device (K2b) instead of np.linalg.inv (:626,:660).  ``Sequential.predict``
fuses every LSTM layer and the Dense top into ONE persistent launch.
"""
from __future__ import annotations

import warnings
from typing import List, Optional

import os

import numpy as np
import torch

from . import _cabi as C
from .layers import (Dense, Handle, HoyerRegularizer, InputLayer, LSTM, LSTMCell, OrthogonalRegularizer,
                     ReducedLSTMCell, SingularLSTM, SingularLSTMCell, TimeDistributed, _losses_of)


# --------------------------------------------------------------------------------------------------
# device linear algebra wrappers (K2 / K2b)
# --------------------------------------------------------------------------------------------------
def svd_batched(A, compute_uv=True, return_sweeps=False, check_convergence=True):
    """Batched thin SVD on device.  A: (batch, m, n) or (m, n).  Returns (U, S, Vt) like
    np.linalg.svd(full_matrices=False) -- U (..,m,k), S (..,k) descending, Vt (..,k,n).  ``check_convergence`` reads the
    per-matrix sweep counts back (one small D2H) and warns if a matrix hit the sweep cap (count reported negative)."""
    a = C.dev_tensor(A)
    squeeze = a.dim() == 2
    if squeeze:
        a = a.unsqueeze(0)
    if a.dim() != 3:
        raise ValueError("svd_batched expects (m,n) or (batch,m,n)")
    batch, m, n = (int(s) for s in a.shape)
    k = min(m, n)
    dev = a.device
    S = torch.empty((batch, k), dtype=torch.float32, device=dev)
    U = torch.empty((batch, m, k), dtype=torch.float32, device=dev) if compute_uv else None
    Vt = torch.empty((batch, k, n), dtype=torch.float32, device=dev) if compute_uv else None
    sw = torch.zeros(batch, dtype=torch.int32, device=dev)
    C.check(C.lib().svdlstm_svd_jacobi_batched(C.ptr(a), batch, m, n, C.ptr(U), C.ptr(S), C.ptr(Vt), C.ptr(sw),
                                               C.cur_stream()))
    C.add_launches(1)
    if check_convergence:
        bad = torch.nonzero(sw < 0).reshape(-1)
        if bad.numel():
            warnings.warn("svd_batched: Jacobi SVD hit the sweep cap without converging for matrix index %s of a (%d, %d, %d) "
                          "batch; its factors are not orthogonal to working precision" % (bad.cpu().tolist(), batch, m, n))
    if squeeze:
        S = S[0]
        U = U[0] if U is not None else None
        Vt = Vt[0] if Vt is not None else None
    out = (U, S, Vt) if compute_uv else S
    if return_sweeps:
        return out, sw
    return out


def reduce_factors_batched(triples):
    """[(U_r (m,r), S_r (r,), V_r (r,n)), ...] -> [(B (m,r), C (r,n-r), pivot_ratio (1,)), ...] from ONE launch of K2b:
    B = (U_r * S_r) @ V1, C = inv(V1) @ V2 with V_r = [V1 | V2] (svd_classes_v3.py:622-626, 656-660)."""
    import ctypes
    items = (C.ReduceItem * len(triples))()
    keep, outs = [], []
    for i, (U_r, S_r, V_r) in enumerate(triples):
        U_r = C.dev_tensor(U_r)
        S_r = C.dev_tensor(S_r).reshape(-1)
        V_r = C.dev_tensor(V_r)
        m, r = int(U_r.shape[0]), int(U_r.shape[1])
        n = int(V_r.shape[1])
        if int(V_r.shape[0]) != r or int(S_r.numel()) != r:
            raise ValueError("reduce_factors: inconsistent ranks")
        if r == 0:
            raise ValueError("reduce_factors: every singular value was pruned (rank 0)")
        if U_r.stride(1) != 1 or V_r.stride(1) != 1:
            U_r, V_r = U_r.contiguous(), V_r.contiguous()
        dev = U_r.device
        B = torch.empty((m, r), dtype=torch.float32, device=dev)
        Cm = torch.empty((r, n - r), dtype=torch.float32, device=dev)
        pr = torch.zeros(1, dtype=torch.float32, device=dev)
        items[i] = C.ReduceItem(U_r.data_ptr(), S_r.data_ptr(), V_r.data_ptr(), B.data_ptr(), Cm.data_ptr() if Cm.numel() else None,
                                pr.data_ptr(), U_r.stride(0), V_r.stride(0), m, r, n)
        keep.append((U_r, S_r, V_r))
        outs.append((B, Cm, pr))
    C.check(C.lib().svdlstm_reduce_factors_batched(items, len(triples), C.cur_stream()))
    C.add_launches(1)
    return outs


def reduce_factors(U_r, S_r, V_r, return_pivot_ratio=False):
    """B = (U_r * S_r) @ V1, C = inv(V1) @ V2 with V_r = [V1 | V2] (svd_classes_v3.py:622-626)."""
    B, Cm, pr = reduce_factors_batched([(U_r, S_r, V_r)])[0]
    if return_pivot_ratio:
        return B, Cm, pr
    return B, Cm


# --------------------------------------------------------------------------------------------------
# Sequential
# --------------------------------------------------------------------------------------------------
class Sequential:
    """keras.models.Sequential stand-in: InputLayer + LSTM layers + Dense top."""

    def __init__(self, layers=None, engine=None):
        self._input: Optional[InputLayer] = None
        self.layers: List = []
        self.engine = engine
        self._fused: Optional[Handle] = None
        for l in (layers or []):
            self.add(l)

    def add(self, layer):
        if isinstance(layer, InputLayer):
            self._input = layer
        else:
            self.layers.append(layer)
        self._fused = None

    @property
    def input_shape(self):
        if self._input is not None and self._input.input_shape is not None:
            return (None,) + tuple(self._input.input_shape)
        first = self.layers[0]
        d = first.cell.input_dim if getattr(first, "cell", None) is not None else None
        return (None, None, d)

    def _lstm_layers(self):
        return [l for l in self.layers if isinstance(l, SingularLSTM)]

    def _dense(self):
        last = self.layers[-1] if self.layers else None
        if isinstance(last, TimeDistributed):
            return last.layer
        if isinstance(last, Dense):
            return last
        return None

    def build(self, input_shape=None):
        d = (input_shape or self.input_shape)[-1]
        if d is None:
            raise ValueError("Sequential needs an InputLayer(input_shape=[None, D]) or an explicit input_shape")
        for l in self._lstm_layers():
            l.build((None, None, d))
            d = l.units
        dense = self._dense()
        if dense is not None and not dense.built:
            dense.build((None, d))

    def get_weights(self):
        out = []
        for l in self.layers:
            out += l.get_weights()
        return out

    def count_params(self):
        return int(sum(np.size(w) for w in self.get_weights()))

    @property
    def losses(self):
        """All regulariser penalties of the model from ONE fused K3 launch (Keras `model.losses`)."""
        items = []
        for l in self._lstm_layers():
            items += l.cell.regularization_items()
        return _losses_of(items)

    def _fusable(self):
        lstms = self._lstm_layers()
        if not lstms or len(lstms) > 8:
            return False
        n_tail = len(self.layers) - len(lstms)
        if n_tail > 1 or (n_tail == 1 and self._dense() is None):
            return False
        if self.layers[:len(lstms)] != lstms:
            return False
        for i, l in enumerate(lstms):
            last = i == len(lstms) - 1
            if l.go_backwards or l.stateful or l.time_major or l.return_state:
                return False
            if not last and not l.return_sequences:
                return False
        dense = self._dense()
        if dense is not None and dense.units > 64:
            return False
        return True

    def _fused_handle(self) -> Handle:
        if self._fused is None:
            lstms = self._lstm_layers()
            h = Handle(lstms[0].cell.input_dim, [l.units for l in lstms])
            for i, l in enumerate(lstms):
                l.cell.bind(h, i)
            dense = self._dense()
            if dense is not None:
                h.set_dense_top(dense.kernel.tensor, dense.bias.tensor)
                import weakref
                href = weakref.ref(h)

                def _dense_changed(dense=dense):
                    # Dense.set_weights / Variable.assign update the buffers in place; the tensor-core engine bakes the Dense
                    # kernel into its FP16 weight-stream image, so the handle has to be told (marks that image stale)
                    hh = href()
                    if hh is None:
                        return False
                    hh.set_dense_top(dense.kernel.tensor, dense.bias.tensor)
                    return True
                dense.kernel._listeners.append(_dense_changed)
                dense.bias._listeners.append(_dense_changed)
            self._fused = h
        return self._fused

    def __call__(self, X, engine=None):
        """Device-in / device-out forward (torch CUDA tensors)."""
        x = C.dev_tensor(X)
        if x.dim() != 3:
            raise ValueError("expected input of shape (batch, time, features)")
        self.build((None, None, int(x.shape[-1])))
        eng = engine if engine is not None else self.engine
        if self._fusable():
            lstms = self._lstm_layers()
            y, _, _ = self._fused_handle().forward(x, return_sequences=lstms[-1].return_sequences, engine=eng)
            return y
        a = x
        for l in self.layers:
            a = l.call(a, engine=eng) if isinstance(l, SingularLSTM) else l.call(a)
        return a

    def stream(self, X, state=None, engine=None):
        """Stateful streaming on the device (the reference's ``stateful=True`` use, svd_classes_v3.py:421-426, for the
        real-time setting of svd_acceleration_v3.py:151): run one time chunk X (batch, time, features) starting from
        ``state`` = (hs, cs) -- lists with one (batch, units) tensor per LSTM layer, as returned by the previous call, or
        None for zeros -- and return (y, state).  Chunked runs equal one long run: bit for bit on the tensor-core
        engine, to float32 rounding on the FP32 engines."""
        x = C.dev_tensor(X)
        if x.dim() != 3:
            raise ValueError("expected input of shape (batch, time, features)")
        self.build((None, None, int(x.shape[-1])))
        if not self._fusable():
            raise ValueError("stream() needs a stack of LSTM layers (+ Dense top) that runs as one fused handle")
        eng = engine if engine is not None else self.engine
        y, hs, cs = self._fused_handle().forward(x, initial_state=state, return_sequences=True, want_state=True, engine=eng)
        return y, (hs, cs)

    def predict(self, X, batch_size=32, verbose=0, engine=None):
        """Keras ``model.predict`` (svd_acceleration_v3.py:148,151): host array in, host array out.
        Sequences are independent, so the whole batch runs as one launch regardless of batch_size.
        Host input (numpy, or a torch CPU tensor -- pinned memory makes the copy asynchronous and full-speed) is staged through
        the serving pipeline of ``predict_async``; a CUDA tensor skips the input copy."""
        return self.predict_async(X, engine=engine, _own_result=True).result()   # an array the caller owns (pooled pinned memory)

    def predict_async(self, X, engine=None, _own_result=False):
        """Enqueue one ``predict`` and return at once with a handle whose ``.result()`` blocks for the host array.

        A serving loop keeps two requests in flight: the host->device copy of request i+1 (own copy stream, second device
        buffer) overlaps the forward pass of request i, and each result is copied into one of two pinned host buffers on
        a third stream (under the forward pass of request i+1).  ``.result()`` returns a numpy VIEW of that pinned buffer, valid until the second-next request
        of the same shape is issued (copy it to keep it longer).  A host input must stay unmodified until ``.result()``
        returns (its copy to the device is asynchronous when the buffer is pinned)."""
        dev = C.require_cuda()
        if isinstance(X, torch.Tensor) and X.is_cuda:
            y = self.__call__(X, engine=engine)
            return _Pending(None, y, None)
        return self._predict_host(X, engine, _own_result)

    def _predict_host(self, X, engine, own_result):
        dev = C.require_cuda()
        xh = X if isinstance(X, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(X, dtype=np.float32)))
        if xh.dtype != torch.float32 or not xh.is_contiguous():
            xh = xh.to(torch.float32).contiguous()
        if xh.dim() != 3:
            raise ValueError("expected input of shape (batch, time, features)")
        key = (tuple(xh.shape), dev.index)
        st = getattr(self, "_serve", None)
        if st is None or st["key"] != key:
            st = {"key": key, "i": 0, "x_dev": [torch.empty(xh.shape, dtype=torch.float32, device=dev) for _ in range(2)],
                  "y_host": [None, None], "copy_stream": torch.cuda.Stream(device=dev), "d2h_stream": torch.cuda.Stream(device=dev),
                  "fwd_done": [torch.cuda.Event() for _ in range(2)],
                  "h2d_done": [torch.cuda.Event() for _ in range(2)], "x_free": [torch.cuda.Event() for _ in range(2)],
                  "y_done": [torch.cuda.Event() for _ in range(2)]}
            main = torch.cuda.current_stream(dev)
            for e in st["x_free"]:
                e.record(main)
            self._serve = st
        k = st["i"] & 1
        st["i"] += 1
        main = torch.cuda.current_stream(dev)
        y = None
        st["copy_stream"].wait_event(st["x_free"][k])              # the forward that last read this buffer has finished
        eng = engine if engine is not None else self.engine
        B, T = int(xh.shape[0]), int(xh.shape[1])
        # Nothing else in flight (a single blocking predict): put the upload inside the forward.  With a request already running the
        # plain contiguous copy is the better upload -- it hides under that forward anyway, and sliced 2-D copies reach only ~35 of
        # the ~55 GB/s a contiguous one does.
        idle = st["i"] < 2 or st["y_done"][1 - k].query()
        if (idle and xh.is_pinned() and eng in (None, "auto", "tc", "tc_f16") and B >= C.TC_MIN_BATCH and T >= 64
                and os.environ.get("SVDLSTM_STREAMED_INPUT", "1") != "0"):
            self.build((None, None, int(xh.shape[-1])))
            if self._fusable() and self._lstm_layers()[-1].return_sequences:
                # the upload goes INSIDE the forward: time slices on the copy stream, the kernel follows them (a single blocking
                # predict then costs max(upload, forward) instead of their sum); None = not a launch that can do it
                y = self._fused_handle().forward_streamed_input(xh, st["x_dev"][k], st["copy_stream"],
                                                               n_slices=int(os.environ.get("SVDLSTM_INPUT_SLICES", "32")))
                if y is None:
                    st["h2d_done"][k].record(st["copy_stream"])        # the slices were enqueued all the same
                    main.wait_event(st["h2d_done"][k])
                    y = self.__call__(st["x_dev"][k], engine=engine)
                elif y is False:
                    y = None                                           # not a tensor-core model: nothing was uploaded yet
        if y is None:
            with torch.cuda.stream(st["copy_stream"]):
                st["x_dev"][k].copy_(xh, non_blocking=True)
                st["h2d_done"][k].record(st["copy_stream"])
            main.wait_event(st["h2d_done"][k])
            y = self.__call__(st["x_dev"][k], engine=engine)
        st["x_free"][k].record(main)
        if own_result:
            y_np = C.pooled_pinned_array(tuple(y.shape))       # the caller keeps it: no copy out of a recycled staging buffer
            y_host = torch.from_numpy(y_np)
        else:
            if st["y_host"][k] is None or tuple(st["y_host"][k].shape) != tuple(y.shape):
                st["y_host"][k] = torch.empty(tuple(y.shape), dtype=torch.float32).pin_memory()
            y_host, y_np = st["y_host"][k], None
        # the result goes home on its own stream: the next request's forward starts right behind this one instead of behind its copy
        st["fwd_done"][k].record(main)
        d2h = st["d2h_stream"]
        with torch.cuda.stream(d2h):
            d2h.wait_event(st["fwd_done"][k])
            y_host.copy_(y, non_blocking=True)
            y.record_stream(d2h)
            st["y_done"][k].record(d2h)
        return _Pending(st["y_done"][k], None, y_host, y_np)

    # ---- training (svd_acceleration_v3.py:111-128) ------------------------------------------------------------------------
    def compile(self, loss="mse", optimizer="adam", learning_rate=None, **kwargs):
        """keras ``model.compile``: only what the reference uses -- loss "mse", optimizer "adam" (Keras defaults:
        learning_rate 1e-3, beta_1 0.9, beta_2 0.999, epsilon 1e-7; ``optimizer`` may also be a dict of those)."""
        from .training import Trainer
        if loss not in ("mse", "mean_squared_error"):
            raise NotImplementedError("loss=%r: the reference trains with 'mse'" % (loss,))
        hp = {}
        if isinstance(optimizer, dict):
            hp = {k: v for k, v in optimizer.items() if k in ("learning_rate", "beta_1", "beta_2", "epsilon")}
        elif str(optimizer).lower() != "adam":
            raise NotImplementedError("optimizer=%r: the reference trains with 'adam'" % (optimizer,))
        if learning_rate is not None:
            hp["learning_rate"] = learning_rate
        self.build()
        self._trainer = Trainer(self, **hp)

    def fit(self, x=None, y=None, batch_size=32, epochs=1, verbose=0, validation_data=None, shuffle=True, **kwargs):
        from .training import fit
        return fit(self, x, y, batch_size=batch_size, epochs=epochs, validation_data=validation_data, shuffle=shuffle, verbose=verbose,
                   seed=kwargs.get("seed"), steps_per_epoch=kwargs.get("steps_per_epoch"))

    def evaluate(self, x=None, y=None, batch_size=256, verbose=0):
        from .training import evaluate
        return evaluate(self, x, y, batch_size=batch_size)

    def train_on_batch(self, x, y):
        if getattr(self, "_trainer", None) is None:
            raise RuntimeError("You must compile your model before training/testing. Use `model.compile(optimizer, loss)`.")
        self._trainer.loss_and_grad(x, y)
        self._trainer.apply()
        return float(self._trainer.loss[0])

    def gradients(self, x, y, regularizers=True):
        """(total loss, [d loss / d w for w in get_weights()]) of one batch -- the quantity `fit` feeds to Adam."""
        if getattr(self, "_trainer", None) is None:
            raise RuntimeError("You must compile your model before training/testing. Use `model.compile(optimizer, loss)`.")
        self._trainer.loss_and_grad(x, y, with_regs=regularizers)
        return float(self._trainer.loss[0]), self._trainer.grads_as_weights()

    def open_stream(self, idle_ms=50):
        """Real-time batch-1 service (the reference's deployment: ``model.predict`` per 16-sample frame every 400-500 us on a
        stateful LSTM; svd_classes_v3.py:421-426, train_full_model_v4.py:14-16).  Launches ONE persistent kernel and returns a
        ``RealtimeStream``; each ``step(x_t)`` then costs no CUDA call (host-mapped rings, see include/svdlstm.h)."""
        self.build()
        if not self._fusable():
            raise ValueError("open_stream() needs a stack of LSTM layers (+ Dense top) that runs as one fused handle")
        return RealtimeStream(self._fused_handle(), idle_ms)

    def last_engine(self):
        return self._fused.last_engine() if self._fused is not None else None


class RealtimeStream:
    """Handle of one persistent real-time kernel (C-ABI ``svdlstm_stream_*``).  ``step`` = one sample in, one prediction out,
    state carried on the device; ``run`` feeds a whole series paced at ``period_us`` from native code and returns the
    per-sample latencies.  Use as a context manager or call ``close()``."""

    def __init__(self, handle: Handle, idle_ms=50):
        import ctypes
        self._handle = handle        # keeps the weights alive
        self._s = ctypes.c_void_p()
        C.require_cuda()
        C.check(C.lib().svdlstm_stream_open(handle.raw, int(idle_ms), ctypes.byref(self._s)))
        self.input_dim = handle.input_dim
        self.n_out = handle.n_out if handle.n_out > 0 else handle.units[-1]
        self._x = np.zeros(self.input_dim, np.float32)
        self._y = np.zeros(self.n_out, np.float32)
        self._launches_seen = 0

    def step(self, x_t) -> np.ndarray:
        np.copyto(self._x, np.asarray(x_t, np.float32).reshape(-1))
        C.check(C.lib().svdlstm_stream_step(self._s, self._x.ctypes.data, self._y.ctypes.data))
        self._count_launches()
        return self._y.copy()

    def run(self, X, period_us=0.0):
        """X (n, input_dim) host array -> (Y (n, n_out), latency_us (n,)): every sample is written when it is due
        (i * period_us after the first) and its latency is the time until its prediction is visible on the host."""
        X = np.ascontiguousarray(np.asarray(X, np.float32).reshape(-1, self.input_dim))
        n = int(X.shape[0])
        Y = np.empty((n, self.n_out), np.float32)
        lat = np.empty(n, np.float32)
        C.check(C.lib().svdlstm_stream_run(self._s, X.ctypes.data, n, float(period_us), Y.ctypes.data, lat.ctypes.data))
        self._count_launches()
        return Y, lat

    def reset_states(self, states=None):
        """Zero state, or (hs, cs): one (units,) / (1, units) array per layer each (``SingularLSTM.reset_states`` semantics)."""
        if states is None:
            C.check(C.lib().svdlstm_stream_reset(self._s, None, None))
            return
        hs, cs = states
        h = np.ascontiguousarray(np.concatenate([np.asarray(_host(a), np.float32).reshape(-1) for a in hs]))
        c = np.ascontiguousarray(np.concatenate([np.asarray(_host(a), np.float32).reshape(-1) for a in cs]))
        if h.size != sum(self._handle.units) or c.size != h.size:
            raise ValueError("state needs one (units,) vector per layer for h and for c")
        C.check(C.lib().svdlstm_stream_reset(self._s, h.ctypes.data, c.ctypes.data))

    def states(self):
        tot = sum(self._handle.units)
        h, c = np.empty(tot, np.float32), np.empty(tot, np.float32)
        C.check(C.lib().svdlstm_stream_state(self._s, h.ctypes.data, c.ctypes.data))
        hs, cs, off = [], [], 0
        for u in self._handle.units:
            hs.append(h[off:off + u].copy())
            cs.append(c[off:off + u].copy())
            off += u
        return hs, cs

    def kernel_launches(self) -> int:
        """How many times the persistent kernel has been (re)launched: 1 for an uninterrupted stream."""
        return int(C.lib().svdlstm_stream_launches(self._s))

    def _count_launches(self):
        n = self.kernel_launches()
        C.add_launches(n - self._launches_seen)
        self._launches_seen = n

    def close(self):
        if getattr(self, "_s", None) is not None and self._s.value:
            C.check(C.lib().svdlstm_stream_close(self._s))
            self._s.value = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _host(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a


class _Pending:
    """Handle of one in-flight ``predict_async`` request."""

    def __init__(self, event, y_dev, y_host, y_np=None):
        self._event, self._y_dev, self._y_host, self._y_np = event, y_dev, y_host, y_np

    def result(self) -> np.ndarray:
        if self._y_host is None:
            return self._y_dev.cpu().numpy()
        self._event.synchronize()
        if self._y_np is not None:          # predict(): pooled pinned memory owned by the caller from here on
            y, self._y_np, self._y_host = self._y_np, None, None
            return y
        return self._y_host.numpy()


# --------------------------------------------------------------------------------------------------
# builders
# --------------------------------------------------------------------------------------------------
def _regs(hoyer, orthogonal):
    if hoyer is not None and hoyer != 0:
        kr, rr = HoyerRegularizer(hoyer), HoyerRegularizer(hoyer)
    else:
        kr = rr = None
    if orthogonal is not None and orthogonal != 0:
        uvr, train_uv = OrthogonalRegularizer(factor=orthogonal, mode='rows'), True
    else:
        uvr, train_uv = None, False
    return kr, rr, uvr, train_uv


def _full_weights(layer):
    if not isinstance(layer.cell, LSTMCell):
        raise ValueError("builder expects a model of stock LSTM layers (get_weights() -> [W,U,b])")
    if not layer.built:
        raise ValueError("source model is not built")
    return layer.cell.kernel.tensor, layer.cell.recurrent_kernel.tensor, layer.cell.bias.tensor


def _copy_dense_top(model, smodel, time_distributed):
    src = model._dense()
    if src is None:
        raise ValueError("builder expects the last layer of the source model to be Dense")
    dense_top = Dense(src.units)
    top = TimeDistributed(dense_top) if time_distributed else dense_top
    smodel.add(top)
    top.set_weights([model.layers[-1].weights[0].numpy(), model.layers[-1].weights[1].numpy()])


def make_split_LSTM_singular_model(model, hoyer=None, orthogonal=None, return_sequences=False):
    """svd_classes_v3.py:469-540: one SVD per gate block i,f,c,o of W and U (batched on device),
    factors re-concatenated along axis 1.  (The reference drops `orthogonal` on this path, :552;
    it is forwarded here.)"""
    model.build()
    smodel = Sequential()
    smodel.add(InputLayer(input_shape=[None, model.input_shape[-1]]))
    lstm_layers = model.layers[:-1]
    for i, layer in enumerate(lstm_layers):
        w, u, b = _full_weights(layer)
        units = layer.units
        wu = []
        for mat in (w, u):
            rows = int(mat.shape[0])
            blocks = mat.view(rows, 4, units).permute(1, 0, 2).contiguous()     # (4, rows, units)
            left, sigma, right = svd_batched(blocks)                            # (4,rows,k) (4,k) (4,k,units)
            k = int(sigma.shape[1])
            unsplit_left = left.permute(1, 0, 2).reshape(rows, 4 * k)
            unsplit_sigma = sigma.reshape(1, 4 * k)
            unsplit_right = right.permute(1, 0, 2).reshape(k, 4 * units)
            wu.append([unsplit_left, unsplit_sigma, unsplit_right])
        kr, rr, uvr, train_uv = _regs(hoyer, orthogonal)
        cell = SingularLSTMCell(units, w=wu[0], u=wu[1], b=b, kernel_regularizer=kr, recurrent_regularizer=rr,
                                train_uv=train_uv, uv_regularizer=uvr, merged_kernel=False)
        rs = True
        if i == len(lstm_layers) - 1 and not return_sequences:
            rs = False
        smodel.add(SingularLSTM(units, cell=cell, return_sequences=rs))
    _copy_dense_top(model, smodel, return_sequences)
    smodel.build()
    return smodel


def make_LSTM_singular_model(model, hoyer=None, orthogonal=None, merged_kernel=True, return_sequences=False):
    """svd_classes_v3.py:548-598: W (D,4H) and U (H,4H) of every LSTM layer -> left, sigma (1,k), right."""
    if not merged_kernel:
        return make_split_LSTM_singular_model(model, hoyer=hoyer, orthogonal=orthogonal, return_sequences=return_sequences)
    model.build()
    smodel = Sequential()
    smodel.add(InputLayer(input_shape=[None, model.input_shape[-1]]))
    lstm_layers = model.layers[:-1]
    for i, layer in enumerate(lstm_layers):
        w, u, b = _full_weights(layer)
        units = layer.units
        wu = []
        for mat in (w, u):
            left, sigma, right = svd_batched(mat)
            wu.append([left, sigma.reshape(1, -1), right])
        kr, rr, uvr, train_uv = _regs(hoyer, orthogonal)
        cell = SingularLSTMCell(units, w=wu[0], u=wu[1], b=b, kernel_regularizer=kr, recurrent_regularizer=rr,
                                train_uv=train_uv, uv_regularizer=uvr)
        rs = True
        if i == len(lstm_layers) - 1 and not return_sequences:
            rs = False
        smodel.add(SingularLSTM(units, cell=cell, return_sequences=rs))
    _copy_dense_top(model, smodel, return_sequences)
    smodel.build()
    return smodel


def _select_factors(U, S, V, cutoff, rank, use_abs=False, where=""):
    """svd_classes_v3.py:618-621.  Threshold keep = S > cutoff (drops negative sigma regardless of
    magnitude, as the reference does; ``use_abs`` opts into |S| > cutoff) or explicit top-``rank``."""
    S = S.reshape(-1)
    if rank is not None:
        r = min(int(rank), S.numel())
        if r < 1:
            raise ValueError("make_LSTM_reduced_model: rank must be >= 1")
        return U[:, :r], S[:r], V[:r]                 # leading factors: strided views, no gather
    s_host = S.cpu()
    keep = (s_host.abs() if use_abs else s_host) > cutoff
    keep_idx = torch.nonzero(keep).reshape(-1).to(S.device)
    if keep_idx.numel() == 0:
        raise ValueError("make_LSTM_reduced_model: cutoff removed every singular value of %s" % where)
    return U.index_select(1, keep_idx).contiguous(), S.index_select(0, keep_idx).contiguous(), V.index_select(0, keep_idx).contiguous()


def make_LSTM_reduced_model(model, cutoff=.05, merged_kernel=True, *, rank=None, use_abs=False, check_condition=True):
    """svd_classes_v3.py:604-676.  From a model of SingularLSTMCells build the 2-factor model.
    ``rank=`` (keyword-only extension) keeps the top-``rank`` factors of every matrix instead of
    thresholding.  Output layers always return sequences + TimeDistributed(Dense) (:630,:665,:670).
    Every (B, C) pair of the model comes out of ONE K2b launch."""
    names, triples, plan = [], [], []
    for li, layer in enumerate(model.layers[:-1]):
        cell = layer.cell
        if not isinstance(cell, SingularLSTMCell):
            raise ValueError("make_LSTM_reduced_model expects a model of SingularLSTMCells")
        if bool(cell.merged_kernel) != bool(merged_kernel):
            raise ValueError("merged_kernel=%s but layer %d holds a %s SingularLSTMCell"
                             % (merged_kernel, li, "merged" if cell.merged_kernel else "split"))
        w_s, u_s, w_l, w_r, u_l, u_r, b = [v.tensor for v in cell.weights]
        if merged_kernel:
            for nm, mat in (("W", (w_l, w_s, w_r)), ("U", (u_l, u_s, u_r))):
                names.append("layer %d %s" % (li, nm))
                triples.append(_select_factors(mat[0], mat[1], mat[2], cutoff, rank, use_abs, names[-1]))
        else:
            w_l4, w_s4, w_r4 = (torch.chunk(a, 4, dim=1) for a in (w_l, w_s, w_r))
            u_l4, u_s4, u_r4 = (torch.chunk(a, 4, dim=1) for a in (u_l, u_s, u_r))
            for g in range(4):
                names.append("layer %d W gate %d" % (li, g))
                triples.append(_select_factors(w_l4[g], w_s4[g], w_r4[g], cutoff, rank, use_abs, names[-1]))
                names.append("layer %d U gate %d" % (li, g))
                triples.append(_select_factors(u_l4[g], u_s4[g], u_r4[g], cutoff, rank, use_abs, names[-1]))
        plan.append((layer.units, b))
    outs = reduce_factors_batched(triples)
    rmodel = Sequential()
    rmodel.add(InputLayer(input_shape=[None, model.input_shape[-1]]))
    per = 2 if merged_kernel else 8
    for li, (units, b) in enumerate(plan):
        o = outs[li * per:(li + 1) * per]
        if merged_kernel:
            rcell = ReducedLSTMCell(units, w=[o[0][0], o[0][1]], u=[o[1][0], o[1][1]], b=b)
        else:
            w = [[o[2 * g][0], o[2 * g][1]] for g in range(4)]
            u = [[o[2 * g + 1][0], o[2 * g + 1][1]] for g in range(4)]
            rcell = ReducedLSTMCell(units, w=w, u=u, b=b, merged_kernel=False)
        rmodel.add(SingularLSTM(units, cell=rcell, return_sequences=True))
    _copy_dense_top(model, rmodel, True)
    rmodel.build()
    if check_condition and outs:
        ratios = torch.cat([o[2] for o in outs]).cpu().numpy()
        rmodel.pivot_ratios = {name: float(r) for name, r in zip(names, ratios)}
        bad = [n for n, r in rmodel.pivot_ratios.items() if not (r > 1e-6)]
        if bad:
            warnings.warn("make_LSTM_reduced_model: V1 is (near-)singular for %s; the reference calls "
                          "np.linalg.inv unguarded here (svd_classes_v3.py:626). Use the 3-factor model "
                          "for these matrices." % ", ".join(bad))
    return rmodel


def truncate_singular_model(model, rank):
    """Top-r truncation of a 3-factor model (keep first r columns / entries / rows of every factor;
    SURVEY App. A) -- the explicit-rank sweep primitive for the 3-factor form.  ``rank`` is one rank for every
    matrix or a pair (rank_w, rank_u) for the input / recurrent factors (the old API's per-matrix ranks,
    old_versions/svd_classes.py:139-182)."""
    rank_w_req, rank_u_req = (rank if isinstance(rank, (tuple, list)) else (rank, rank))
    smodel = Sequential()
    smodel.add(InputLayer(input_shape=[None, model.input_shape[-1]]))
    for layer in model.layers[:-1]:
        cell = layer.cell
        s_w, s_u, w_l, w_r, u_l, u_r, b = [v.tensor for v in cell.weights]
        H = layer.units
        if cell.merged_kernel:
            rw, ru = min(int(rank_w_req), cell.rank_w), min(int(rank_u_req), cell.rank_u)
            w = [w_l[:, :rw], s_w[:, :rw], w_r[:rw]]
            u = [u_l[:, :ru], s_u[:, :ru], u_r[:ru]]
        else:
            def cut(l, s, r_, k, rank=None):
                kk = min(int(rank), k)
                ls = [l[:, g * k:g * k + kk] for g in range(4)]
                ss = [s[:, g * k:g * k + kk] for g in range(4)]
                return [torch.cat(ls, 1), torch.cat(ss, 1), r_[:kk]]
            w = cut(w_l, s_w, w_r, cell.rank_w, rank_w_req)
            u = cut(u_l, s_u, u_r, cell.rank_u, rank_u_req)
        ncell = SingularLSTMCell(H, w=w, u=u, b=b, merged_kernel=cell.merged_kernel,
                                 kernel_regularizer=cell.kernel_regularizer, recurrent_regularizer=cell.recurrent_regularizer,
                                 train_uv=cell.train_uv, uv_regularizer=cell.uv_regularizer)
        smodel.add(SingularLSTM(H, cell=ncell, return_sequences=layer.return_sequences))
    td = isinstance(model.layers[-1], TimeDistributed)
    _copy_dense_top(model, smodel, td)
    smodel.build()
    return smodel


def full_model_from_weights(layers, dense, return_sequences=True):
    """[(W (D,4H), U (H,4H), b (4H,)), ...] + (dense_kernel (H,n), dense_bias (n,)) -> Sequential of
    stock LSTM layers + Dense: the stand-in for keras.models.load_model (svd_acceleration_v3.py:115)."""
    m = Sequential()
    m.add(InputLayer(input_shape=[None, int(np.shape(layers[0][0])[0])]))
    for i, (W, U, b) in enumerate(layers):
        H = int(np.shape(U)[0])
        rs = True if i < len(layers) - 1 else return_sequences
        m.add(LSTM(H, weights=[W, U, b], return_sequences=rs))
    dk, db = dense
    d = Dense(int(np.shape(dk)[1]))
    top = TimeDistributed(d) if return_sequences else d
    m.add(top)
    top.set_weights([dk, db])
    m.build()
    return m
