"""Host-side mirror of the reference's operator API (code/svd_classes_v3.py:17-465).

Same class names, constructor kwargs, weight orderings and call contracts as the reference's Keras
subclasses; every forward pass goes through the C-ABI (``_cabi``) into hand-written sm_100a CUDA.
Nothing here computes on the CPU.

    SingularLSTMCell   svd_classes_v3.py:17-236   3-factor cell   get_weights() = [sigma_w, sigma_u,
                                                   w_left, w_right, u_left, u_right, bias] (:113)
    ReducedLSTMCell    svd_classes_v3.py:240-368  2-factor cell   [w_left,w_right,u_left,u_right,bias] or
                                                   per-gate x4 + bias (:278,:308-315)
    LSTMCell           stock keras LSTMCell       [W, U, b] (what the builders read at :557)
    SingularLSTM       svd_classes_v3.py:375-440  layer / time loop (backend.rnn) around a cell
    LSTM               keras.layers.LSTM stand-in (full cell)
    HoyerRegularizer   svd_classes_v3.py:455-465
    OrthogonalRegularizer  keras.regularizers.OrthogonalRegularizer(mode='rows'), call sites :514,:573
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _cabi as C

_default_engine = "auto"


def set_default_engine(name: str) -> None:
    """'auto' | 'fp32' | 'general' | 'wavefront' | 'tc' -- engine used when a call does not name one.  'auto' is the regime
    switch of include/svdlstm.h (tensor cores for dense batches, FP32 otherwise); 'fp32' never leaves full precision."""
    global _default_engine
    if name not in C.ENGINE_NAMES:
        raise ValueError("unknown engine %r" % (name,))
    _default_engine = name


def get_default_engine() -> str:
    return _default_engine


def _engine_id(engine) -> int:
    if engine is None:
        engine = _default_engine
    if isinstance(engine, int):
        return engine
    if engine not in C.ENGINE_NAMES:
        raise ValueError("unknown engine %r" % (engine,))
    return C.ENGINE_NAMES[engine]


class Variable:
    """Stand-in for tf.Variable: ``.numpy()``, ``.shape``, ``.name``; ``.tensor`` is the device buffer."""

    def __init__(self, tensor: torch.Tensor, name: str, trainable: bool = False, regularizer=None):
        self.tensor = tensor
        self.name = name
        self.trainable = trainable
        self.regularizer = regularizer
        # callables run after every in-place update: whoever bound this buffer into a device handle registers one, so that
        # derived device state (the packed FP16 weight-stream images of the tensor-core engine) is rebuilt on the next forward
        self._listeners = []

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    def numpy(self) -> np.ndarray:
        return self.tensor.detach().cpu().numpy()

    def assign(self, value, notify: bool = True) -> None:
        v = C.dev_tensor(value, self.tensor.device)
        if tuple(v.shape) != tuple(self.tensor.shape):
            raise ValueError("Layer weight shape %s not compatible with provided weight shape %s"
                             % (tuple(self.tensor.shape), tuple(v.shape)))
        self.tensor.copy_(v)
        if notify:
            self.notify()

    def notify(self) -> None:
        alive = []
        for fn in self._listeners:
            if fn() is not False:       # a listener whose target died returns False and is dropped
                alive.append(fn)
        self._listeners = alive

    def __repr__(self):
        return "<Variable %s shape=%s>" % (self.name, self.shape)


class Handle:
    """Owns one svdlstm_handle and keeps the bound device tensors alive."""

    def __init__(self, input_dim: int, units: Sequence[int]):
        self._h = ctypes.c_void_p()
        self.units = [int(u) for u in units]
        self.input_dim = int(input_dim)
        C.check(C.lib().svdlstm_create(ctypes.byref(self._h), len(self.units), self.input_dim,
                                       C.int_array(self.units)))
        self._keep = {}
        self.n_out = 0

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                C.lib().svdlstm_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass

    @property
    def raw(self):
        return self._h

    def set_dense_top(self, kernel: Optional[torch.Tensor], bias: Optional[torch.Tensor]):
        if kernel is None:
            C.check(C.lib().svdlstm_set_dense_top(self._h, None, None, 0))
            self.n_out = 0
            self._keep.pop("dense", None)
            return
        if kernel.dim() != 2 or kernel.shape[0] != self.units[-1] or bias.numel() != kernel.shape[1]:
            raise ValueError("Dense top expects kernel (%d, n_out) and bias (n_out,), got %s / %s"
                             % (self.units[-1], tuple(kernel.shape), tuple(bias.shape)))
        self._keep["dense"] = (kernel, bias)
        self.n_out = int(kernel.shape[1])
        C.check(C.lib().svdlstm_set_dense_top(self._h, C.ptr(kernel), C.ptr(bias), self.n_out))

    def forward(self, x: torch.Tensor, *, initial_state=None, mask=None, return_sequences=True,
                go_backwards=False, time_major=False, zero_output_for_mask=False, want_state=False,
                engine=None):
        """x: (B,T,D) float32 CUDA (or (T,B,D) with time_major).  Returns (y, h_n, c_n)."""
        dev = x.device
        if x.dim() != 3 or x.shape[-1] != self.input_dim:
            raise ValueError("expected input of shape (batch, time, %d), got %s" % (self.input_dim, tuple(x.shape)))
        if time_major:
            T, B = int(x.shape[0]), int(x.shape[1])
        else:
            B, T = int(x.shape[0]), int(x.shape[1])
        n = self.n_out if self.n_out > 0 else self.units[-1]
        eid = _engine_id(engine)
        dense_regime = eid == C.ENGINE_TC or (eid == C.ENGINE_AUTO and B >= C.TC_MIN_BATCH and max(self.units) >= C.TC_MIN_UNITS)
        if (time_major or go_backwards) and mask is None and dense_regime:
            # The tensor-core kernel walks a batch-major sequence forwards.  time_major (svd_classes_v3.py:408-419 via backend.rnn)
            # is a transpose of the input and of the output sequence; go_backwards is the time-reversed input, outputs staying in
            # processing order as Keras returns them.  Two cheap passes over x instead of the FP32 engine for the whole forward.
            xs = x.transpose(0, 1) if time_major else x
            if go_backwards:
                xs = xs.flip(1)
            y, hs_out, cs_out = self.forward(xs.contiguous(), initial_state=initial_state, return_sequences=return_sequences,
                                             want_state=want_state, engine=engine)
            if return_sequences and time_major:
                y = y.transpose(0, 1).contiguous()
            return y, hs_out, cs_out
        if (not return_sequences and not time_major and not go_backwards and initial_state is None and mask is None and not want_state
                and dense_regime):
            # The tensor-core kernel always produces the whole output sequence (the Dense top is fused into its S1 tiles, one
            # step behind); return_sequences=False (svd_classes_v3.py:428-431: last output only) is the last step of it.
            try:
                y, _, _ = self.forward(x, return_sequences=True, engine="tc")
                return y[:, -1].contiguous(), None, None
            except ValueError:
                if eid == C.ENGINE_TC:
                    raise
                engine = "fp32"     # "auto", and the tensor-core engine does not take this model: FP32 engines
        if return_sequences:
            y = torch.empty((T, B, n) if time_major else (B, T, n), dtype=torch.float32, device=dev)
        else:
            y = torch.empty((B, n), dtype=torch.float32, device=dev)
        tot = sum(self.units) * B
        h0 = c0 = None
        if initial_state is not None:
            hs, cs = initial_state
            hs = hs if isinstance(hs, (list, tuple)) else [hs]
            cs = cs if isinstance(cs, (list, tuple)) else [cs]
            if len(hs) != len(self.units) or len(cs) != len(self.units):
                raise ValueError("initial_state needs one (h, c) pair per layer")
            for l, (hh, cc) in enumerate(zip(hs, cs)):
                if tuple(hh.shape) != (B, self.units[l]) or tuple(cc.shape) != (B, self.units[l]):
                    raise ValueError("initial state of layer %d must have shape (%d, %d)" % (l, B, self.units[l]))
            h0 = torch.cat([C.dev_tensor(t, dev).reshape(-1) for t in hs])
            c0 = torch.cat([C.dev_tensor(t, dev).reshape(-1) for t in cs])
        h_n = c_n = None
        if want_state:
            h_n = torch.empty(tot, dtype=torch.float32, device=dev)
            c_n = torch.empty(tot, dtype=torch.float32, device=dev)
        m = None
        if mask is not None:
            m = torch.as_tensor(mask, device=dev)
            if time_major:
                m = m.transpose(0, 1)
            if tuple(m.shape) != (B, T):
                raise ValueError("mask must have shape (batch, time)")
            m = (m != 0).to(torch.uint8).contiguous()
        flags = ((C.RETURN_SEQUENCES if return_sequences else 0) | (C.GO_BACKWARDS if go_backwards else 0)
                 | (C.TIME_MAJOR if time_major else 0) | (C.ZERO_OUTPUT_FOR_MASK if zero_output_for_mask else 0))
        L = C.lib()
        C.check(L.svdlstm_forward(self._h, C.ptr(x), B, T, C.ptr(y), C.ptr(h0), C.ptr(c0), C.ptr(h_n), C.ptr(c_n),
                                  C.ptr(m), flags, _engine_id(engine), C.cur_stream()))
        C.add_launches(L.svdlstm_last_launches(self._h))
        if want_state:
            hs_out, cs_out, off = [], [], 0
            for u in self.units:
                hs_out.append(h_n[off:off + B * u].view(B, u))
                cs_out.append(c_n[off:off + B * u].view(B, u))
                off += B * u
            return y, hs_out, cs_out
        return y, None, None

    def forward_streamed_input(self, x_host: torch.Tensor, x_dev: torch.Tensor, copy_stream, n_slices: int = 0):
        """Forward with the host->device upload of ``x_host`` (pinned, float32, (B,T,D)) INSIDE it: the array goes up in time
        slices on ``copy_stream`` and the tensor-core kernel follows the upload (``svdlstm_forward_streamed_input``).  Returns
        the output sequence; ``None`` if the launch this batch takes cannot follow an upload -- the slices are enqueued on
        ``copy_stream`` all the same, so the caller waits for it and calls :meth:`forward` on ``x_dev``; ``False`` if the
        tensor-core engine does not take the model at all (nothing was enqueued)."""
        B, T = int(x_host.shape[0]), int(x_host.shape[1])
        n = self.n_out if self.n_out > 0 else self.units[-1]
        y = torch.empty((B, T, n), dtype=torch.float32, device=x_dev.device)
        L = C.lib()
        rc = L.svdlstm_forward_streamed_input(self._h, x_host.data_ptr(), C.ptr(x_dev), B, T, C.ptr(y), int(n_slices),
                                              copy_stream.cuda_stream, C.cur_stream())
        if rc == -3:
            return False
        if rc == -4:
            return None
        C.check(rc)
        C.add_launches(L.svdlstm_last_launches(self._h))
        return y

    def last_engine(self) -> int:
        return C.lib().svdlstm_last_engine(self._h)

    def count_weights(self) -> int:
        return int(C.lib().svdlstm_count_weights(self._h))


# --------------------------------------------------------------------------------------------------
# regularizers
# --------------------------------------------------------------------------------------------------
def evaluate_penalties(items) -> np.ndarray:
    """ONE fused K3 launch for a list of (tensor, gram: bool, columns: bool).  Returns (n,4) float64:
    [sum|x|, sum x^2, sum_{i!=j}|Pn_ij|, ||G - I||_F^2]."""
    dev = C.require_cuda()
    arr = (C.PenaltyItem * len(items))()
    keep = []
    for i, (t, gram, columns) in enumerate(items):
        t = C.dev_tensor(t, dev)
        if t.dim() == 1:
            t = t.view(1, -1)
        if t.dim() != 2:
            raise ValueError("penalty items must be vectors or rank-2 matrices")
        keep.append(t)
        arr[i] = C.PenaltyItem(t.data_ptr(), t.shape[0], t.shape[1], t.stride(0), 1 if gram else 0, 1 if columns else 0)
    out = torch.empty(4 * len(items), dtype=torch.float64, device=dev)
    C.check(C.lib().svdlstm_penalties(arr, len(items), C.ptr(out), C.cur_stream()))
    C.add_launches(1)
    return out.cpu().numpy().reshape(len(items), 4)


class HoyerRegularizer:
    """svd_classes_v3.py:455-465.  ``__call__`` = hoyer * sum|x| / sum x^2 (the reference formula, no
    square root); ``l1_over_l2`` is the north-star's sum|s|/||s||_2 variant from the same raw sums."""

    gram = False
    columns = False

    def __init__(self, hoyer=0.):
        hoyer = 0 if hoyer is None else hoyer
        self.hoyer = np.float32(hoyer)

    def from_raw(self, raw, shape=None) -> float:
        return float(self.hoyer) * float(raw[0]) / float(raw[1])

    def l1_over_l2(self, x) -> float:
        raw = evaluate_penalties([(_as_tensor(x), False, False)])[0]
        return float(raw[0]) / float(np.sqrt(raw[1]))

    def __call__(self, x) -> float:
        return self.from_raw(evaluate_penalties([(_as_tensor(x), False, False)])[0])

    def get_config(self):
        return {'hoyer': self.hoyer}


class OrthogonalRegularizer:
    """keras.regularizers.OrthogonalRegularizer(factor, mode) as used at svd_classes_v3.py:514,573:
    factor * 0.5 * sum|P o (1-I)| / (n(n-1)/2) with P the Gram matrix of the L2-normalised rows
    (mode='rows') or columns.  ``fro_sq`` gives the north-star's ||X X^T - I||_F^2."""

    gram = True

    def __init__(self, factor=0.01, mode='rows'):
        if mode not in ('rows', 'columns'):
            raise ValueError("Invalid value for argument `mode`. Expected one of {'rows', 'columns'}. Received: mode=%s" % mode)
        self.factor = float(factor)
        self.mode = mode

    @property
    def columns(self):
        return self.mode == 'columns'

    def from_raw(self, raw, shape) -> float:
        n = shape[1] if self.columns else shape[0]
        num_pairs = n * (n - 1.0) / 2.0
        if num_pairs == 0:
            return 0.0
        return self.factor * 0.5 * float(raw[2]) / num_pairs

    def __call__(self, x) -> float:
        t = _as_tensor(x)
        if t.dim() != 2:
            raise ValueError("Inputs to OrthogonalRegularizer must have rank 2. Received: inputs.shape == %s" % (tuple(t.shape),))
        return self.from_raw(evaluate_penalties([(t, True, self.columns)])[0], tuple(t.shape))

    def fro_sq(self, x) -> float:
        t = _as_tensor(x)
        return float(evaluate_penalties([(t, True, self.columns)])[0][3])

    def get_config(self):
        return {'factor': self.factor, 'mode': self.mode}


def _as_tensor(x):
    if isinstance(x, Variable):
        return x.tensor
    return C.dev_tensor(x)


# --------------------------------------------------------------------------------------------------
# cells
# --------------------------------------------------------------------------------------------------
_UNSUPPORTED_CELL_KW = {"activation": "tanh", "recurrent_activation": "sigmoid", "use_bias": True,
                        "dropout": 0.0, "recurrent_dropout": 0.0}


class _CellBase:
    merged_kernel = True

    def __init__(self, units, **kwargs):
        self.units = int(units)
        for k, default in _UNSUPPORTED_CELL_KW.items():
            if k in kwargs and kwargs[k] not in (default, None) and kwargs[k] != default:
                raise NotImplementedError("%s=%r: only the Keras LSTMCell defaults (%r) are implemented on device"
                                          % (k, kwargs[k], default))
        self.name = kwargs.get("name", type(self).__name__.lower())
        self.built = False
        self._vars: List[Variable] = []
        self._handle: Optional[Handle] = None
        self.input_dim: Optional[int] = None

    # Keras protocol ----------------------------------------------------------------------------
    @property
    def state_size(self):
        return [self.units, self.units]

    @property
    def output_size(self):
        return self.units

    @property
    def weights(self) -> List[Variable]:
        return list(self._vars)

    @property
    def trainable_weights(self):
        return [v for v in self._vars if v.trainable]

    @property
    def non_trainable_weights(self):
        return [v for v in self._vars if not v.trainable]

    def get_weights(self):
        self._require_built()
        return [v.numpy() for v in self._vars]

    def set_weights(self, weights):
        self._require_built()
        if len(weights) != len(self._vars):
            raise ValueError('You called `set_weights(weights)` on layer "%s" with a weight list of length %d, '
                             'but the layer was expecting %d weights.' % (self.name, len(weights), len(self._vars)))
        for v, w in zip(self._vars, weights):
            if tuple(np.shape(w)) != v.shape:
                raise ValueError("Layer weight shape %s not compatible with provided weight shape %s"
                                 % (v.shape, tuple(np.shape(w))))
        for v, w in zip(self._vars, weights):
            v.assign(w, notify=False)
        self.rebind()

    def count_params(self):
        self._require_built()
        return int(sum(int(np.prod(v.shape)) for v in self._vars))

    def _require_built(self):
        if not self.built:
            raise ValueError("Weights for cell %s have not yet been created. Call build(input_shape) or the cell first." % self.name)

    def _add(self, value, shape, name, trainable=False, regularizer=None) -> Variable:
        t = C.dev_tensor(value)
        if isinstance(value, (torch.Tensor, Variable)):
            # Every cell OWNS its buffers, as Keras variables do: the builders pass live tensors / views of the source model
            # (its bias, slices of its factors), and an in-place set_weights on one model must never change another.
            t = t.clone()
        if tuple(t.shape) != tuple(shape):
            raise ValueError("Layer weight shape %s not compatible with provided weight shape %s (weight %r of %s)"
                             % (tuple(shape), tuple(t.shape), name, self.name))
        v = Variable(t, name, trainable, regularizer)
        import weakref
        me = weakref.ref(self)

        def _on_assign():     # Variable.assign outside set_weights (e.g. cell.kernel.assign(...)): re-announce the layer to its handles
            c = me()
            if c is None:
                return False
            c.rebind()
            return True
        v._listeners.append(_on_assign)
        self._vars.append(v)
        return v

    def regularization_items(self):
        """[(regularizer, Variable)] for every weight that carries one (Keras `losses`)."""
        return [(v.regularizer, v) for v in self._vars if v.regularizer is not None]

    @property
    def losses(self):
        return _losses_of(self.regularization_items())

    # device binding ----------------------------------------------------------------------------
    def bind(self, handle: Handle, layer: int) -> None:
        """Hands this cell's device buffers to layer `layer` of `handle` (borrowed pointers)."""
        self._bind(handle, layer)
        import weakref
        if not hasattr(self, "_bound"):
            self._bound = []
        if not any(r() is handle and l == layer for r, l in self._bound):
            self._bound.append((weakref.ref(handle), layer))

    def rebind(self) -> None:
        """After an in-place weight update: tell every handle (packed tensor-core copies are stale)."""
        for r, l in list(getattr(self, "_bound", [])):
            h = r()
            if h is not None:
                self._bind(h, l)

    def _bind(self, handle: Handle, layer: int) -> None:
        raise NotImplementedError

    def _own_handle(self) -> Handle:
        if self._handle is None:
            self._handle = Handle(self.input_dim, [self.units])
            self.bind(self._handle, 0)
        return self._handle

    def call(self, inputs, states, training=None):
        """One timestep: inputs (B,D), states [h (B,H), c (B,H)] -> (h, [h, c])
        (svd_classes_v3.py:116,236,317,368)."""
        x = C.dev_tensor(inputs)
        if not self.built:
            self.build(tuple(x.shape))
        h_tm1, c_tm1 = states[0], states[1]
        y, hs, cs = self._own_handle().forward(x.unsqueeze(1), initial_state=([C.dev_tensor(h_tm1)], [C.dev_tensor(c_tm1)]),
                                               return_sequences=False, want_state=True, engine="general")
        h, c = hs[0], cs[0]
        return h, [h, c]

    __call__ = call

    def get_initial_state(self, inputs=None, batch_size=None, dtype=None):
        b = batch_size if batch_size is not None else int(inputs.shape[0])
        dev = C.require_cuda()
        return [torch.zeros((b, self.units), dtype=torch.float32, device=dev) for _ in range(2)]


def _losses_of(items):
    if not items:
        return []
    raws = evaluate_penalties([(v.tensor, r.gram, r.columns) for r, v in items])
    return [r.from_raw(raw, v.shape) for (r, v), raw in zip(items, raws)]


class LSTMCell(_CellBase):
    """Stock Keras LSTMCell maths; weights [kernel (D,4H), recurrent_kernel (H,4H), bias (4H,)]."""

    def __init__(self, units, w=None, u=None, b=None, **kwargs):
        super().__init__(units, **kwargs)
        self.w, self.u, self.b = w, u, b

    def build(self, input_shape):
        if self.built:
            return
        D = int(input_shape[-1])
        H = self.units
        if self.w is None or self.u is None or self.b is None:
            raise ValueError("LSTMCell needs trained weights w, u, b (no initialisers: training is out of scope)")
        self.input_dim = D
        self.kernel = self._add(self.w, (D, 4 * H), "kernel", True)
        self.recurrent_kernel = self._add(self.u, (H, 4 * H), "recurrent_kernel", True)
        self.bias = self._add(self.b, (4 * H,), "bias", True)
        self.built = True

    def _bind(self, handle, layer):
        C.check(C.lib().svdlstm_set_full_weights(handle.raw, layer, C.ptr(self.kernel.tensor),
                                                 C.ptr(self.recurrent_kernel.tensor), C.ptr(self.bias.tensor)))
        handle._keep[layer] = [v.tensor for v in self._vars]


class SingularLSTMCell(_CellBase):
    """svd_classes_v3.py:17-236.  ``w = [left, sigma (1,k), right]``, ``u`` likewise.

    Deviations from the reference as written (built to evident intent):
    merged ``build`` works (the reference has a ``regularzier=`` typo, :54); ranks are read from the
    passed factors (k = sigma.shape[1], per gate k/4 when split) instead of being hard-coded to
    input_dim / units (:36,:59,:67,:75,:96) so that D > H and truncated factors are accepted."""

    def __init__(self, units, w=None, u=None, b=None, merged_kernel=True,
                 train_uv=False, kernel_regularizer=None,
                 recurrent_regularizer=None, uv_regularizer=None, **kwargs):
        super().__init__(units, **kwargs)
        self.w = w; self.u = u; self.b = b
        self.train_uv = train_uv
        self.kernel_regularizer = kernel_regularizer
        self.recurrent_regularizer = recurrent_regularizer
        self.merged_kernel = merged_kernel
        self.uv_regularizer = uv_regularizer

    def build(self, input_shape):
        if self.built:
            return
        if self.w is None or self.u is None or self.b is None:
            raise ValueError("SingularLSTMCell needs factors w=[left,sigma,right], u=[left,sigma,right] and bias b")
        D = int(input_shape[-1])
        H = self.units
        self.input_dim = D
        kw_tot = int(np.shape(self.w[1])[-1])
        ku_tot = int(np.shape(self.u[1])[-1])
        if self.merged_kernel:
            kw, ku = kw_tot, ku_tot
            if kw > min(D, 4 * H) or ku > H:
                raise ValueError("merged ranks (%d,%d) exceed min(D,4H)=%d / H=%d" % (kw, ku, min(D, 4 * H), H))
        else:
            if kw_tot % 4 or ku_tot % 4:
                raise ValueError("split kernel / recurrent_kernel must hold 4 per-gate blocks")
            kw, ku = kw_tot // 4, ku_tot // 4
        self.rank_w, self.rank_u = kw, ku
        tv = self.train_uv
        # creation order == get_weights() order (svd_classes_v3.py:35-113)
        self.kernel = self._add(np.reshape(_np(self.w[1]), (1, kw_tot)), (1, kw_tot), "kernel", True, self.kernel_regularizer)
        self.recurrent_kernel = self._add(np.reshape(_np(self.u[1]), (1, ku_tot)), (1, ku_tot), "recurrent_kernel", True,
                                          self.recurrent_regularizer)
        self.w_left = self._add(self.w[0], (D, kw_tot), "w_left", tv, self.uv_regularizer)
        self.w_right = self._add(self.w[2], (kw, 4 * H), "w_right", tv, self.uv_regularizer)
        self.u_left = self._add(self.u[0], (H, ku_tot), "u_left", tv, self.uv_regularizer)
        self.u_right = self._add(self.u[2], (ku, 4 * H), "u_right", tv, self.uv_regularizer)
        self.bias = self._add(self.b, (4 * H,), "bias", tv)
        self.built = True

    def _bind(self, handle, layer):
        ts = [v.tensor for v in self._vars]
        C.check(C.lib().svdlstm_set_singular_weights(handle.raw, layer, 1 if self.merged_kernel else 0,
                                                     C.ptr_array(ts), self.rank_w, self.rank_u))
        handle._keep[layer] = ts


class ReducedLSTMCell(_CellBase):
    """svd_classes_v3.py:240-368.  merged: ``w=[B,C]``, ``u=[B,C]``; split: ``w=[[B,C]]*4``."""

    def __init__(self, units, w=None, u=None, b=None, merged_kernel=True, **kwargs):
        super().__init__(units, **kwargs)
        self.w = w; self.u = u; self.b = b
        self.merged_kernel = merged_kernel

    def build(self, input_shape):
        if self.built:
            return
        if self.w is None or self.u is None or self.b is None:
            raise ValueError("ReducedLSTMCell needs factors w, u and bias b")
        D = int(input_shape[-1])
        H = self.units
        self.input_dim = D
        if self.merged_kernel:
            rank_w = int(np.shape(self.w[0])[1])
            rank_u = int(np.shape(self.u[0])[1])
            self.ranks = [rank_w, rank_u]
            self.w_left = self._add(self.w[0], (D, rank_w), "w_left")
            self.w_right = self._add(self.w[1], (rank_w, H * 4 - rank_w), "w_right")
            self.u_left = self._add(self.u[0], (H, rank_u), "u_left")
            self.u_right = self._add(self.u[1], (rank_u, H * 4 - rank_u), "u_right")
        else:
            self.w_left = []; self.w_right = []
            self.u_left = []; self.u_right = []
            self.ranks = []
            for i in range(4):
                gate = ['i', 'f', 'c', 'o'][i]
                rank_w = int(np.shape(self.w[i][0])[1])
                rank_u = int(np.shape(self.u[i][0])[1])
                self.ranks += [rank_w, rank_u]
                self.w_left.append(self._add(self.w[i][0], (D, rank_w), "w_left_" + gate))
                self.w_right.append(self._add(self.w[i][1], (rank_w, H - rank_w), "w_right_" + gate))
                self.u_left.append(self._add(self.u[i][0], (H, rank_u), "u_left_" + gate))
                self.u_right.append(self._add(self.u[i][1], (rank_u, H - rank_u), "u_right_" + gate))
        self.bias = self._add(self.b, (4 * H,), "bias")
        self.built = True

    def _bind(self, handle, layer):
        ts = [v.tensor for v in self._vars]
        # zero-sized right factors (rank == full width) are passed as NULL
        ptrs = [t if t.numel() > 0 else None for t in ts]
        C.check(C.lib().svdlstm_set_reduced_weights(handle.raw, layer, 1 if self.merged_kernel else 0,
                                                    C.ptr_array(ptrs), C.int_array(self.ranks)))
        handle._keep[layer] = ts


def _np(a):
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    if isinstance(a, Variable):
        return a.numpy()
    return np.asarray(a)


# --------------------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------------------
class SingularLSTM:
    """svd_classes_v3.py:375-440: keras.layers.LSTM with an injected cell; ``call`` runs the generic
    time loop (backend.rnn, :408-419) -- here ONE persistent CUDA launch for all T steps."""

    def __init__(self, units, cell=None, return_sequences=False, return_state=False, go_backwards=False,
                 stateful=False, time_major=False, unroll=False, zero_output_for_mask=False, name=None,
                 engine=None, **kwargs):
        self.units = int(units)
        self.cell = cell
        self.return_sequences = return_sequences
        self.return_state = return_state
        self.go_backwards = go_backwards
        self.stateful = stateful
        self.time_major = time_major
        self.unroll = unroll            # accepted; meaningless for a persistent kernel
        self.zero_output_for_mask = zero_output_for_mask
        self.name = name or type(self).__name__.lower()
        self.engine = engine
        self.states = None
        self._handle: Optional[Handle] = None
        if cell is not None and cell.units != self.units:
            raise ValueError("cell.units (%d) != layer units (%d)" % (cell.units, self.units))

    # Keras-ish surface ---------------------------------------------------------------------------
    @property
    def built(self):
        return self.cell is not None and self.cell.built

    def build(self, input_shape):
        if self.cell is None:
            raise ValueError("%s needs a cell (the reference always injects one, svd_classes_v3.py:381)" % self.name)
        self.cell.build(input_shape)

    @property
    def weights(self):
        return self.cell.weights

    @property
    def trainable_weights(self):
        return self.cell.trainable_weights

    def get_weights(self):
        return self.cell.get_weights()

    def set_weights(self, weights):
        self.cell.set_weights(weights)

    def count_params(self):
        return self.cell.count_params()

    def get_prunable_weights(self):
        return [self.cell.kernel, self.cell.recurrent_kernel]     # svd_classes_v3.py:439-440

    @property
    def losses(self):
        return self.cell.losses

    def reset_states(self, states=None):
        self.states = None if states is None else [C.dev_tensor(s) for s in states]

    def _own_handle(self) -> Handle:
        if self._handle is None:
            self._handle = Handle(self.cell.input_dim, [self.units])
            self.cell.bind(self._handle, 0)
        return self._handle

    def call(self, inputs, mask=None, training=None, initial_state=None, engine=None):
        x = C.dev_tensor(inputs)
        if not self.built:
            self.build(tuple(x.shape))
        if isinstance(mask, list):
            mask = mask[0]
        if initial_state is None and self.stateful and self.states is not None:
            initial_state = self.states
        init = None
        if initial_state is not None:
            init = ([initial_state[0]], [initial_state[1]])
        want_state = self.return_state or self.stateful
        y, hs, cs = self._own_handle().forward(
            x, initial_state=init, mask=mask, return_sequences=self.return_sequences, go_backwards=self.go_backwards,
            time_major=self.time_major, zero_output_for_mask=self.zero_output_for_mask, want_state=want_state,
            engine=engine if engine is not None else self.engine)
        if self.stateful:
            self.states = [hs[0].clone(), cs[0].clone()]      # svd_classes_v3.py:421-426
        if self.return_state:
            return [y, hs[0], cs[0]]
        return y

    __call__ = call


class LSTM(SingularLSTM):
    """Stand-in for the stock keras.layers.LSTM of the trained full model (what the builders consume:
    ``layer.get_weights() -> [W,U,b]``, ``layer.units``; svd_classes_v3.py:474,557)."""

    def __init__(self, units, weights=None, cell=None, **kwargs):
        if cell is None:
            if weights is None:
                raise ValueError("LSTM needs trained weights=[W,U,b] (training is out of scope)")
            cell = LSTMCell(units, w=weights[0], u=weights[1], b=weights[2])
        super().__init__(units, cell=cell, **kwargs)


class Dense:
    """keras.layers.Dense (linear).  Inside a Sequential the Dense top is fused into the recurrent
    kernel; called on its own it runs through K5 (svdlstm_scaled_matmul)."""

    def __init__(self, units, name=None):
        self.units = int(units)
        self.name = name or "dense"
        self.built = False
        self._vars: List[Variable] = []

    def build(self, input_shape):
        if self.built:
            return
        dev = C.require_cuda()
        d = int(input_shape[-1])
        self.kernel = Variable(torch.zeros((d, self.units), dtype=torch.float32, device=dev), "kernel", True)
        self.bias = Variable(torch.zeros((self.units,), dtype=torch.float32, device=dev), "bias", True)
        self._vars = [self.kernel, self.bias]
        self.built = True

    @property
    def weights(self):
        return list(self._vars)

    def get_weights(self):
        return [v.numpy() for v in self._vars]

    def set_weights(self, weights):
        if not self.built:
            self.build((None, int(np.shape(weights[0])[0])))
        if len(weights) != 2:
            raise ValueError("Dense expects [kernel, bias]")
        for v, w in zip(self._vars, weights):
            w = _np(w)
            if w.size != int(np.prod(v.shape)):
                raise ValueError("Layer weight shape %s not compatible with provided weight shape %s" % (v.shape, w.shape))
            v.assign(w.reshape(v.shape))

    def count_params(self):
        return int(sum(int(np.prod(v.shape)) for v in self._vars))

    def call(self, inputs):
        x = C.dev_tensor(inputs)
        if not self.built:
            raise ValueError("Dense has no weights yet")
        lead = tuple(x.shape[:-1])
        x2 = x.reshape(-1, x.shape[-1])
        return C.scaled_matmul(x2, self.kernel.tensor, bias=self.bias.tensor).reshape(lead + (self.units,))

    __call__ = call


class TimeDistributed:
    """keras.layers.TimeDistributed(Dense): the same Dense applied to every timestep."""

    def __init__(self, layer, name=None):
        self.layer = layer
        self.name = name or "time_distributed"

    @property
    def built(self):
        return self.layer.built

    @property
    def weights(self):
        return self.layer.weights

    def get_weights(self):
        return self.layer.get_weights()

    def set_weights(self, weights):
        self.layer.set_weights(weights)

    def count_params(self):
        return self.layer.count_params()

    def call(self, inputs):
        return self.layer.call(inputs)

    __call__ = call


class PrunableTimeDistributed(TimeDistributed):
    """svd_classes_v3.py:442-449 (tfmot hook; pruning itself is out of scope)."""

    def get_prunable_weights(self):
        return self.layer.weights


class InputLayer:
    def __init__(self, input_shape=None, name=None):
        self.input_shape = tuple(input_shape) if input_shape is not None else None
        self.name = name or "input"
