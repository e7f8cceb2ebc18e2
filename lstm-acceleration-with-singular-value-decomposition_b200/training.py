"""compile / fit for models of SingularLSTMCells: the Hoyer fine-tune of the reference driver on device.

    smodel.compile(loss="mse", optimizer="adam")                       code/svd_acceleration_v3.py:111-118
    smodel.fit(X_mini, y_mini, batch_size=32, validation_data=..., epochs=10)      :119-128

Trainable weights as in the reference: ``kernel`` / ``recurrent_kernel`` (the singular values) always, the factor
matrices and the bias only with ``train_uv`` (svd_classes_v3.py:40-56, 102-112), the Dense top always.  Regularisers
(HoyerRegularizer on the sigma vectors, :455-462; OrthogonalRegularizer(mode='rows') on the factors, :514,:573) enter
the loss and its gradient exactly as Keras adds ``layer.losses``.  Everything runs through the C-ABI (K6:
``svdlstm_trainer_*``): forward with cache + back-propagation through time, regulariser gradients, Adam in place.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _cabi as C
from .layers import HoyerRegularizer, OrthogonalRegularizer, SingularLSTMCell


class History:
    """keras.callbacks.History stand-in: ``.history`` = {"loss": [...], "val_loss": [...]} per epoch."""

    def __init__(self):
        self.history = {}
        self.epoch = []


class Trainer:
    """Owns one svdlstm_trainer bound to the model's fused handle."""

    def __init__(self, model, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        lstms = model._lstm_layers()
        if not model._fusable() or model._dense() is None:
            raise ValueError("compile()/fit() need a stack of SingularLSTM layers + a Dense top")
        for l in lstms:
            if not isinstance(l.cell, SingularLSTMCell):
                raise ValueError("only models of SingularLSTMCells (make_LSTM_singular_model) are trainable on device")
        self.model = model
        self.return_sequences = bool(lstms[-1].return_sequences)
        self.handle = model._fused_handle()
        self.hp = (float(learning_rate), float(beta_1), float(beta_2), float(epsilon))
        self._t = ctypes.c_void_p()
        C.check(C.lib().svdlstm_trainer_create(self.handle.raw, C.int_array([1 if l.cell.train_uv else 0 for l in lstms]),
                                               ctypes.byref(self._t)))
        self.n_params = int(C.lib().svdlstm_trainer_num_params(self._t))
        L = len(lstms)
        offs = (ctypes.c_int64 * (7 * L + 3))()
        C.check(C.lib().svdlstm_trainer_layout(self._t, offs))
        self.offsets = [int(o) for o in offs]
        dev = C.require_cuda()
        self.grad = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        # regulariser work list: (tensor index 7 l + w, kind, coefficient)
        idx, kind, coef = [], [], []
        for li, l in enumerate(lstms):
            c = l.cell
            for w, reg in ((0, c.kernel_regularizer), (1, c.recurrent_regularizer)):
                self._add_reg(idx, kind, coef, 7 * li + w, reg)
            if c.train_uv:      # non-trainable weights carry no loss in Keras
                for w in (2, 3, 4, 5):
                    self._add_reg(idx, kind, coef, 7 * li + w, c.uv_regularizer)
        self.regs = (C.int_array(idx), C.int_array(kind), (ctypes.c_float * len(coef))(*coef), len(idx))

    @staticmethod
    def _add_reg(idx, kind, coef, tensor_index, reg):
        if reg is None:
            return
        if isinstance(reg, HoyerRegularizer):
            if float(reg.hoyer) != 0.0:
                idx.append(tensor_index); kind.append(1); coef.append(float(reg.hoyer))
        elif isinstance(reg, OrthogonalRegularizer):
            if reg.mode != "rows":
                raise NotImplementedError("OrthogonalRegularizer(mode='columns') has no device gradient (the reference uses 'rows')")
            if reg.factor != 0.0:
                idx.append(tensor_index); kind.append(2); coef.append(float(reg.factor))
        else:
            raise NotImplementedError("regulariser %r has no device gradient" % (reg,))

    def __del__(self):
        try:
            if getattr(self, "_t", None) is not None and self._t.value:
                C.lib().svdlstm_trainer_destroy(self._t)
                self._t = ctypes.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------
    def _targets(self, y, B, T):
        n_y = self.handle.n_out
        yt = C.dev_tensor(y)
        want = (B, T, n_y) if self.return_sequences else (B, n_y)
        if yt.numel() != int(np.prod(want)):
            raise ValueError("targets have %d values, the model outputs %s" % (yt.numel(), want))
        return yt.reshape(want).contiguous()

    def loss_and_grad(self, x, y, with_grad=True, with_regs=True):
        """Enqueues forward + backward (+ regulariser terms).  Afterwards self.loss (device scalar) holds the total loss
        and self.grad the flat gradient."""
        xt = C.dev_tensor(x)
        if xt.dim() != 3 or xt.shape[-1] != self.handle.input_dim:
            raise ValueError("expected input of shape (batch, time, %d)" % self.handle.input_dim)
        B, T = int(xt.shape[0]), int(xt.shape[1])
        yt = self._targets(y, B, T)
        L = C.lib()
        C.check(L.svdlstm_trainer_gradients(self._t, C.ptr(xt), C.ptr(yt), B, T, 1 if self.return_sequences else 0,
                                            C.ptr(self.grad) if with_grad else None, C.ptr(self.loss), C.cur_stream()))
        C.add_launches(2)
        if with_regs and self.regs[3] > 0:
            if not with_grad:
                raise ValueError("regulariser terms are evaluated together with the gradient")
            C.check(L.svdlstm_trainer_regularizers(self._t, self.regs[0], self.regs[1], self.regs[2], self.regs[3],
                                                   C.ptr(self.grad), C.ptr(self.loss), C.cur_stream()))
            C.add_launches(1)

    def apply(self):
        lr, b1, b2, eps = self.hp
        C.check(C.lib().svdlstm_trainer_adam(self._t, C.ptr(self.grad), lr, b1, b2, eps, C.cur_stream()))
        C.add_launches(1)

    def grads_as_weights(self):
        """The flat gradient split like ``model.get_weights()`` (numpy; zeros for non-trainable tensors)."""
        g = self.grad.cpu().numpy()
        out = []
        for li, l in enumerate(self.model._lstm_layers()):
            for w, v in enumerate(l.cell.weights):
                o = self.offsets[7 * li + w]
                out.append(g[o:o + int(np.prod(v.shape))].reshape(v.shape).copy())
        nL = len(self.model._lstm_layers())
        dk, db = self.model._dense().weights
        ok, ob, oe = self.offsets[7 * nL], self.offsets[7 * nL + 1], self.offsets[7 * nL + 2]
        out.append(g[ok:ob].reshape(dk.shape).copy())
        out.append(g[ob:oe].reshape(db.shape).copy())
        return out


def fit(model, X, y, batch_size=32, epochs=1, validation_data=None, shuffle=True, verbose=0, seed=None, steps_per_epoch=None) -> History:
    tr: Optional[Trainer] = getattr(model, "_trainer", None)
    if tr is None:
        raise RuntimeError("You must compile your model before training/testing. Use `model.compile(optimizer, loss)`.")
    dev = C.require_cuda()
    Xd = C.dev_tensor(X)                     # the whole training set lives on the device (reference: 20 000 x 200 x 16 floats = 256 MB)
    N, T = int(Xd.shape[0]), int(Xd.shape[1])
    yd = tr._targets(y, N, T)
    rng = np.random.default_rng(seed)
    hist = History()
    hist.history["loss"] = []
    if validation_data is not None:
        hist.history["val_loss"] = []
    n_batches = (N + batch_size - 1) // batch_size
    if steps_per_epoch is not None:
        n_batches = min(n_batches, int(steps_per_epoch))
    for ep in range(int(epochs)):
        order = torch.as_tensor(rng.permutation(N) if shuffle else np.arange(N), device=dev)
        losses = torch.zeros(n_batches, dtype=torch.float32, device=dev)
        for bi in range(n_batches):
            sel = order[bi * batch_size:(bi + 1) * batch_size]
            tr.loss_and_grad(Xd.index_select(0, sel), yd.index_select(0, sel))
            tr.apply()
            losses[bi] = tr.loss[0]
        ep_loss = float(losses.mean())          # one host read per epoch
        hist.history["loss"].append(ep_loss)
        hist.epoch.append(ep)
        if validation_data is not None:
            hist.history["val_loss"].append(evaluate(model, validation_data[0], validation_data[1], batch_size=max(batch_size, 256)))
        if verbose:
            print("Epoch %d/%d - loss: %.6f%s" % (ep + 1, epochs, ep_loss,
                                                  " - val_loss: %.6f" % hist.history["val_loss"][-1] if validation_data is not None else ""))
    for l in model._lstm_layers():                # in-place updates: announce them to every other handle the cells are bound to
        l.cell.rebind()
    model._dense().kernel.notify()
    model.history = hist
    return hist


def evaluate(model, X, y, batch_size=256) -> float:
    """Mean-squared error over the data set (no regulariser terms, like the `val_loss` data term)."""
    tr: Optional[Trainer] = getattr(model, "_trainer", None)
    if tr is None:
        raise RuntimeError("You must compile your model before training/testing. Use `model.compile(optimizer, loss)`.")
    Xd = C.dev_tensor(X)
    N, T = int(Xd.shape[0]), int(Xd.shape[1])
    n_y = tr.handle.n_out
    y_n = int(np.prod(np.shape(y)))
    if not tr.return_sequences and y_n != N * n_y and y_n % (N * n_y) == 0:
        # The reference validates a last-step model against the WHOLE target series (`validation_data=(X, y.reshape(1, -1, 1))`,
        # svd_acceleration_v3.py:125, with return_sequences=False): Keras broadcasts the (N, n) prediction over the extra axis.
        # Same number here: inference forward + K4 squared error of the broadcast prediction.
        from .metrics import sweep_sse
        pred = model(Xd).reshape(N, 1, n_y)
        yt = C.dev_tensor(y).reshape(N, -1, n_y)
        sse = sweep_sse(pred.expand_as(yt).contiguous().reshape(1, -1), yt.reshape(-1))
        return float(sse[0]) / y_n
    yd = tr._targets(y, N, T)
    tot, cnt = 0.0, 0
    for b0 in range(0, N, batch_size):
        xb, yb = Xd[b0:b0 + batch_size].contiguous(), yd[b0:b0 + batch_size].contiguous()
        tr.loss_and_grad(xb, yb, with_grad=False, with_regs=False)
        tot += float(tr.loss[0]) * int(xb.shape[0])
        cnt += int(xb.shape[0])
    return tot / cnt
