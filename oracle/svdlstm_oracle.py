"""CPU oracle for the SVD-factored LSTM hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain numpy, the arithmetic of the reference's hot path
(``code/svd_classes_v3.py`` + the metric / timing part of
``code/svd_acceleration_v3.py`` + the explicit-rank helpers of
``code/old_versions/svd_classes.py``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; the product package never does (it fails loudly
when the CUDA library is missing instead of falling back to this).

PARITY STATUS
-------------
* The reference cannot be imported here (TensorFlow/Keras 2.10 is not installed,
  no network) and it ships no tests/golden vectors for the LSTM forward pass
  (its only forward golden, ``model_prediction.csv``, needs the missing
  ``preprocessed_DROPBEAR_X.csv``).  The *forward pass* of this oracle is
  therefore pinned only by (i) algebraic identities full == 3-factor@full-rank
  == 2-factor@full-rank, (ii) an independent ``torch.nn.LSTM`` CPU cross-check,
  (iii) the cross-session known-answer vector of SURVEY.md App. D.  In the
  prompt's vocabulary: **forward parity unpinned by reference tests**.
* The SVD / inverse boundary (``np.linalg.svd`` / ``np.linalg.inv``, the very
  calls the reference makes at svd_classes_v3.py:491,562,626,660) *is*
  executable here, so the builders are pinned by the reference's own dependency.
* The metric code is pinned by goldens derived from shipped fixtures
  (RMSE 0.20285040751787883, SNR 12.433968928917704 dB; tests/golden).

All functions take ``dtype`` (np.float64 for the checker, np.float32 for the
"reference-faithful op granularity" CPU baseline).  Row-vector convention
(svd_classes_v3.py:128), gate order i,f,c,o along the 4H axis (:144).
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

GATES = ("i", "f", "c", "o")


# --------------------------------------------------------------------------------------
# activations (Keras LSTMCell defaults: activation=tanh, recurrent_activation=sigmoid)
# --------------------------------------------------------------------------------------
def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _carry_and_output_fused(z, c_tm1):
    """Keras ``LSTMCell._compute_carry_and_output_fused`` [un-vendored; SURVEY App. B].
    i=sig(z0); f=sig(z1); c=f*c_tm1+i*tanh(z2); o=sig(z3)."""
    z0, z1, z2, z3 = z
    i = sigmoid(z0)
    f = sigmoid(z1)
    c = f * c_tm1 + i * np.tanh(z2)
    o = sigmoid(z3)
    return c, o


# --------------------------------------------------------------------------------------
# cells.  Every cell is a small class with ``step(x, h, c) -> (h, c)`` and
# ``get_weights()`` in the reference's order.
# --------------------------------------------------------------------------------------
class FullCell:
    """Stock Keras LSTMCell maths: z = x W + h U + b."""

    def __init__(self, units, W, U, b, dtype=np.float64):
        self.units = int(units)
        self.W = np.asarray(W, dtype)
        self.U = np.asarray(U, dtype)
        self.b = np.asarray(b, dtype)
        self.dtype = dtype

    def step(self, x, h, c):
        z = x @ self.W + h @ self.U + self.b
        z = np.split(z, 4, axis=1)
        c, o = _carry_and_output_fused(z, c)
        h = o * np.tanh(c)
        return h, c

    def get_weights(self):
        return [self.W, self.U, self.b]


class SingularCell:
    """``SingularLSTMCell`` (svd_classes_v3.py:17-236).

    ``w = [left, sigma(1,k), right]`` and ``u`` likewise (constructor order, :19-23);
    ``get_weights()`` order is [kernel(sigma_w), recurrent_kernel(sigma_u), w_left,
    w_right, u_left, u_right, bias] (:113)."""

    def __init__(self, units, w, u, b, merged_kernel=True, dtype=np.float64):
        self.units = int(units)
        self.merged_kernel = bool(merged_kernel)
        self.dtype = dtype
        self.w_left = np.asarray(w[0], dtype)
        self.kernel = np.asarray(w[1], dtype).reshape(1, -1)
        self.w_right = np.asarray(w[2], dtype)
        self.u_left = np.asarray(u[0], dtype)
        self.recurrent_kernel = np.asarray(u[1], dtype).reshape(1, -1)
        self.u_right = np.asarray(u[2], dtype)
        self.bias = np.asarray(b, dtype)

    def get_weights(self):
        return [self.kernel, self.recurrent_kernel, self.w_left, self.w_right,
                self.u_left, self.u_right, self.bias]

    def step(self, inputs, h_tm1, c_tm1):
        if self.merged_kernel:
            # svd_classes_v3.py:129-145
            x = inputs @ self.w_left
            x = x * self.kernel
            x = x @ self.w_right
            x = x + self.bias
            z = h_tm1 @ self.u_left
            z = z * self.recurrent_kernel
            z = z @ self.u_right
            z = z + x
            z = np.split(z, 4, axis=1)
            c, o = _carry_and_output_fused(z, c_tm1)
        else:
            # svd_classes_v3.py:165-232 (per-gate factors are axis-1 quarters)
            wr = np.split(self.w_right, 4, axis=1)
            ws = np.split(self.kernel, 4, axis=1)
            wl = np.split(self.w_left, 4, axis=1)
            ur = np.split(self.u_right, 4, axis=1)
            us = np.split(self.recurrent_kernel, 4, axis=1)
            ul = np.split(self.u_left, 4, axis=1)
            bs = np.split(self.bias, 4, axis=0)
            pre = []
            for g in range(4):
                xg = ((inputs @ wl[g]) * ws[g]) @ wr[g] + bs[g]
                rg = ((h_tm1 @ ul[g]) * us[g]) @ ur[g]
                pre.append(xg + rg)
            i = sigmoid(pre[0])
            f = sigmoid(pre[1])
            c = f * c_tm1 + i * np.tanh(pre[2])
            o = sigmoid(pre[3])
        h = o * np.tanh(c)
        return h, c


class ReducedCell:
    """``ReducedLSTMCell`` (svd_classes_v3.py:240-368).

    merged: ``w=[B,C]``, ``u=[B,C]``; split: ``w=[[B,C]]*4`` per gate i,f,c,o (:278,:308).
    ``get_weights()``: merged [w_left,w_right,u_left,u_right,bias]; split
    [w_left_g,w_right_g,u_left_g,u_right_g]*4 + [bias]."""

    def __init__(self, units, w, u, b, merged_kernel=True, dtype=np.float64):
        self.units = int(units)
        self.merged_kernel = bool(merged_kernel)
        self.dtype = dtype
        if merged_kernel:
            self.w_left = np.asarray(w[0], dtype)
            self.w_right = np.asarray(w[1], dtype)
            self.u_left = np.asarray(u[0], dtype)
            self.u_right = np.asarray(u[1], dtype)
        else:
            self.w_left = [np.asarray(w[g][0], dtype) for g in range(4)]
            self.w_right = [np.asarray(w[g][1], dtype) for g in range(4)]
            self.u_left = [np.asarray(u[g][0], dtype) for g in range(4)]
            self.u_right = [np.asarray(u[g][1], dtype) for g in range(4)]
        self.bias = np.asarray(b, dtype)

    def get_weights(self):
        if self.merged_kernel:
            return [self.w_left, self.w_right, self.u_left, self.u_right, self.bias]
        out = []
        for g in range(4):
            out += [self.w_left[g], self.w_right[g], self.u_left[g], self.u_right[g]]
        out.append(self.bias)
        return out

    def step(self, inputs, h_tm1, c_tm1):
        if self.merged_kernel:
            # svd_classes_v3.py:321-328
            x = inputs @ self.w_left
            x = np.concatenate((x, x @ self.w_right), axis=1)
            x = x + self.bias
            z = h_tm1 @ self.u_left
            z = np.concatenate((z, z @ self.u_right), axis=1)
            z = z + x
            z = np.split(z, 4, axis=1)
        else:
            # svd_classes_v3.py:330-363
            bs = np.split(self.bias, 4, axis=0)
            z = []
            for g in range(4):
                xg = inputs @ self.w_left[g]
                xg = np.concatenate((xg, xg @ self.w_right[g]), axis=1) + bs[g]
                zg = h_tm1 @ self.u_left[g]
                zg = np.concatenate((zg, zg @ self.u_right[g]), axis=1)
                z.append(zg + xg)
        c, o = _carry_and_output_fused(z, c_tm1)
        h = o * np.tanh(c)
        return h, c


# --------------------------------------------------------------------------------------
# layer / time loop  (SingularLSTM.call -> backend.rnn, svd_classes_v3.py:385-437)
# --------------------------------------------------------------------------------------
def rnn_layer(cell, inputs, initial_state=None, mask=None, go_backwards=False,
              return_sequences=False, return_state=False, time_major=False,
              zero_output_for_mask=False):
    """Restates ``backend.rnn`` semantics used at svd_classes_v3.py:408-434.

    inputs (B,T,D) (or (T,B,D) if time_major).  Zero initial state when
    ``initial_state`` is None (:393).  With ``go_backwards`` the input is reversed
    in time and outputs come out in processing order (Keras does not re-reverse).
    With ``mask`` (B,T) a masked step carries the previous state and repeats the
    previous output (zeros before the first valid step, or always zeros when
    ``zero_output_for_mask``)."""
    dtype = cell.dtype
    x = np.asarray(inputs, dtype)
    if time_major:
        x = np.swapaxes(x, 0, 1)
    B, T, _ = x.shape
    H = cell.units
    if initial_state is None:
        h = np.zeros((B, H), dtype)
        c = np.zeros((B, H), dtype)
    else:
        h = np.array(initial_state[0], dtype)
        c = np.array(initial_state[1], dtype)
    order = range(T - 1, -1, -1) if go_backwards else range(T)
    outs = np.zeros((B, T, H), dtype)
    prev_out = np.zeros((B, H), dtype)
    for k, t in enumerate(order):
        h_new, c_new = cell.step(x[:, t, :], h, c)
        if mask is not None:
            m = np.asarray(mask)[:, t].astype(bool)[:, None]
            out = np.where(m, h_new, np.zeros_like(h_new) if zero_output_for_mask else prev_out)
            h = np.where(m, h_new, h)
            c = np.where(m, c_new, c)
        else:
            out = h_new
            h, c = h_new, c_new
        outs[:, k, :] = out
        prev_out = out
    last = prev_out
    output = outs if return_sequences else last
    if time_major and return_sequences:
        output = np.swapaxes(output, 0, 1)
    if return_state:
        return [output, h, c]
    return output


class Model:
    """Stand-in for the Keras Sequential the builders take/return: ``cells`` are the
    LSTM layers (``model.layers[:-1]``), ``dense`` = (kernel (H,n_out), bias (n_out,))
    is ``model.layers[-1]``.  ``return_sequences`` describes the *last* LSTM layer
    (all earlier layers always return sequences, svd_classes_v3.py:526-529)."""

    def __init__(self, cells, dense, return_sequences=True, input_dim=None):
        self.cells = list(cells)
        self.dense = (np.asarray(dense[0]), np.asarray(dense[1]))
        self.return_sequences = bool(return_sequences)
        self.input_dim = input_dim

    def predict(self, X):
        a = np.asarray(X, self.cells[0].dtype)
        for li, cell in enumerate(self.cells):
            last = li == len(self.cells) - 1
            a = rnn_layer(cell, a, return_sequences=(self.return_sequences or not last))
        k = self.dense[0].astype(a.dtype)
        b = self.dense[1].astype(a.dtype)
        return a @ k + b   # Dense(1) / TimeDistributed(Dense(1)), svd_classes_v3.py:590-597,670-675

    def get_weights(self):
        out = []
        for c in self.cells:
            out += c.get_weights()
        out += [self.dense[0], self.dense[1]]
        return out


# --------------------------------------------------------------------------------------
# builders (svd_classes_v3.py:469-676)
# --------------------------------------------------------------------------------------
def factor_merged(mat):
    """np.linalg.svd(mat, full_matrices=False) with sigma expanded to (1,k) (:561-564)."""
    left, sigma, right = np.linalg.svd(mat, full_matrices=False, compute_uv=True)
    return [left, np.expand_dims(sigma, 0), right]


def factor_split(mat, units):
    """Per-gate SVDs re-concatenated along axis 1 (:482-502)."""
    blocks = [mat[:, g * units:(g + 1) * units] for g in range(4)]
    lefts, sigmas, rights = [], [], []
    for blk in blocks:
        l, s, r = np.linalg.svd(blk, full_matrices=False, compute_uv=True)
        lefts.append(l)
        sigmas.append(np.expand_dims(s, 0))
        rights.append(r)
    return [np.concatenate(lefts, 1), np.concatenate(sigmas, 1), np.concatenate(rights, 1)]


def make_LSTM_singular_model(model: Model, merged_kernel=True, return_sequences=False,
                             svd_dtype=np.float32, dtype=np.float64):
    """svd_classes_v3.py:548-598 / :469-540.  The SVD runs in the dtype
    ``get_weights()`` hands it (float32 in the reference)."""
    cells = []
    for cell in model.cells:
        W, U, b = cell.get_weights()
        W = np.asarray(W, svd_dtype)
        U = np.asarray(U, svd_dtype)
        if merged_kernel:
            wu = [factor_merged(W), factor_merged(U)]
        else:
            wu = [factor_split(W, cell.units), factor_split(U, cell.units)]
        cells.append(SingularCell(cell.units, wu[0], wu[1], b, merged_kernel=merged_kernel, dtype=dtype))
    return Model(cells, model.dense, return_sequences=return_sequences, input_dim=model.input_dim)


def _reduce_one(U, S, V, cutoff=None, rank=None):
    """svd_classes_v3.py:618-627.  ``cutoff``: keep S>cutoff (reference); ``rank``:
    keep the first ``rank`` entries (the explicit top-r extension)."""
    S = np.asarray(S).reshape(1, -1)
    if rank is not None:
        keep = np.zeros(S.shape[1], bool)
        keep[:min(int(rank), S.shape[1])] = True
    else:
        keep = (S > cutoff)[0]
    U = U.T[keep].T
    V = V[keep]
    S = S[0][keep]
    r = V.shape[0]
    V1 = V[:, :r]
    V2 = V[:, r:]
    B = (U * S) @ V1
    C = np.linalg.inv(V1) @ V2
    return [B, C]


def make_LSTM_reduced_model(model: Model, cutoff=.05, merged_kernel=True, rank=None,
                            dtype=np.float64, work_dtype=None):
    """svd_classes_v3.py:604-676 (output layers always return sequences, :630,:665)."""
    cells = []
    for cell in model.cells:
        weights = cell.get_weights()
        if work_dtype is not None:
            weights = [np.asarray(w, work_dtype) for w in weights]
        w_s, u_s, w_l, w_r, u_l, u_r, b = weights
        units = cell.units
        if merged_kernel:
            wu = [_reduce_one(w_l, w_s, w_r, cutoff, rank), _reduce_one(u_l, u_s, u_r, cutoff, rank)]
            cells.append(ReducedCell(units, wu[0], wu[1], b, merged_kernel=True, dtype=dtype))
        else:
            w_l4, w_s4, w_r4 = (np.split(a, 4, axis=1) for a in (w_l, w_s, w_r))
            u_l4, u_s4, u_r4 = (np.split(a, 4, axis=1) for a in (u_l, u_s, u_r))
            w, u = [], []
            for g in range(4):
                w.append(_reduce_one(w_l4[g], w_s4[g], w_r4[g], cutoff, rank))
                u.append(_reduce_one(u_l4[g], u_s4[g], u_r4[g], cutoff, rank))
            cells.append(ReducedCell(units, w, u, b, merged_kernel=False, dtype=dtype))
    return Model(cells, model.dense, return_sequences=True, input_dim=model.input_dim)


def truncate_singular_model(model: Model, rank, dtype=np.float64):
    """Top-r truncation of a 3-factor model (keep first r columns/entries/rows, App. A)."""
    cells = []
    for cell in model.cells:
        s_w, s_u, w_l, w_r, u_l, u_r, b = cell.get_weights()
        if cell.merged_kernel:
            rw = min(rank, s_w.shape[1])
            ru = min(rank, s_u.shape[1])
            w = [w_l[:, :rw], s_w[:, :rw], w_r[:rw]]
            u = [u_l[:, :ru], s_u[:, :ru], u_r[:ru]]
        else:
            def cut(l, s, r_):
                k = s.shape[1] // 4
                kk = min(rank, k)
                H = r_.shape[1] // 4
                ls = [l[:, g * k:g * k + kk] for g in range(4)]
                ss = [s[:, g * k:g * k + kk] for g in range(4)]
                rs = [r_[:kk, g * H:(g + 1) * H] for g in range(4)]
                return [np.concatenate(ls, 1), np.concatenate(ss, 1), np.concatenate(rs, 1)]
            w = cut(w_l, s_w, w_r)
            u = cut(u_l, s_u, u_r)
        cells.append(SingularCell(cell.units, w, u, b, merged_kernel=cell.merged_kernel, dtype=dtype))
    return Model(cells, model.dense, return_sequences=model.return_sequences, input_dim=model.input_dim)


# --------------------------------------------------------------------------------------
# regularizers
# --------------------------------------------------------------------------------------
def hoyer_regularizer(x, hoyer):
    """svd_classes_v3.py:461 -- hoyer * sum|x| / sum x^2 (no square root)."""
    x = np.asarray(x)
    return hoyer * np.sum(np.abs(x)) / np.sum(np.square(x))


def hoyer_l1_over_l2(x):
    """north_star's variant: sum|s| / ||s||_2."""
    x = np.asarray(x, np.float64)
    return np.sum(np.abs(x)) / math.sqrt(np.sum(np.square(x)))


def orthogonal_regularizer_rows(X, factor):
    """Keras 2.10 ``OrthogonalRegularizer(factor, mode='rows')`` [un-vendored; App. B]:
    factor * 0.5 * sum|P o (1-I)| / (n(n-1)/2), P = Xn Xn^T, Xn row-L2-normalised
    (l2_normalize: x * rsqrt(max(sum x^2, 1e-12)))."""
    X = np.asarray(X, np.float64)
    n = X.shape[0]
    nrm = np.sqrt(np.maximum(np.sum(X * X, axis=1, keepdims=True), 1e-12))
    Xn = X / nrm
    P = Xn @ Xn.T
    off = np.abs(P * (1.0 - np.eye(n)))
    num_pairs = n * (n - 1.0) / 2.0
    return factor * 0.5 * np.sum(off) / num_pairs


def orthogonality_fro_sq(X, mode="rows"):
    """||X X^T - I||_F^2 (rows) or ||X^T X - I||_F^2 (columns) -- north_star's variant."""
    X = np.asarray(X, np.float64)
    G = X @ X.T if mode == "rows" else X.T @ X
    return float(np.sum((G - np.eye(G.shape[0])) ** 2))


def penalty_raw_sums(X, mode="rows"):
    """The four raw sums the fused penalty kernel emits per item:
    [sum|x|, sum x^2, sum_{i!=j}|Pn_ij| (normalised Gram), ||G-I||_F^2 (raw Gram)]."""
    X = np.asarray(X, np.float64)
    l1 = np.sum(np.abs(X))
    l2 = np.sum(X * X)
    if X.ndim != 2 or min(X.shape) == 0:
        return np.array([l1, l2, 0.0, 0.0])
    Y = X if mode == "rows" else X.T
    n = Y.shape[0]
    nrm = np.sqrt(np.maximum(np.sum(Y * Y, axis=1, keepdims=True), 1e-12))
    Pn = (Y / nrm) @ (Y / nrm).T
    off = np.sum(np.abs(Pn)) - np.sum(np.abs(np.diag(Pn)))
    G = Y @ Y.T
    fro = np.sum((G - np.eye(n)) ** 2)
    return np.array([l1, l2, off, fro])


# --------------------------------------------------------------------------------------
# metrics (svd_acceleration_v3.py:90-100,160-170,187-204)
# --------------------------------------------------------------------------------------
def signaltonoise(signal, noisy_signal, invert=False, dB=True):
    noise = signal - noisy_signal
    a_sig = math.sqrt(np.mean(np.square(signal)))
    a_noise = math.sqrt(np.mean(np.square(noise)))
    snr = (a_sig / a_noise) ** 2 if not invert else (a_noise / a_sig) ** 2
    if not dB:
        return snr
    return 10 * math.log(snr, 10)


def rmse(y_true, y_pred):
    d = np.asarray(y_true, np.float64).ravel() - np.asarray(y_pred, np.float64).ravel()
    return math.sqrt(np.sum(d * d) / d.size)


def reference_rmse(y_true, y_pred, divisor_len):
    """svd_acceleration_v3.py:188 -- sum over all of y, divided by len(y_test)."""
    d = np.asarray(y_true, np.float64).ravel() - np.asarray(y_pred, np.float64).ravel()
    return math.sqrt(np.sum(d * d) / divisor_len)


def count_weights(model) -> int:
    """svd_acceleration_v3.py:160-166 -- every array of get_weights(), Dense included."""
    return int(sum(np.asarray(w).size for w in model.get_weights()))


def full_weight_count(D, H):
    return 4 * (D * H + H * H + H)          # slides 8-9


def reduced_split_weight_count(D, H, rw, ru):
    return 4 * (rw * (D + H - rw) + ru * (2 * H - ru)) + 4 * H


def reduced_merged_weight_count(D, H, rw, ru):
    return rw * (D + 4 * H - rw) + ru * (5 * H - ru) + 4 * H


# --------------------------------------------------------------------------------------
# old explicit-rank API (old_versions/svd_classes.py:9-21,210-232)
# --------------------------------------------------------------------------------------
def reduce_matrix_rank(a, rank):
    u, s, v = np.linalg.svd(a, full_matrices=True, compute_uv=True)
    s = s.copy()
    s[rank:] = 0
    k = s.size
    return (u[:, :k] * s) @ v[:k]


def reduce_two_step(a, rank):
    """Column-vector two-step [b s v ; c b^-1] (old_versions/svd_classes.py:14-21)."""
    u, s, v = np.linalg.svd(a, full_matrices=True, compute_uv=True)
    s = np.diag(s)[:rank, :rank]
    v = v[:rank, :]
    u = u[:, :rank]
    b = u[:rank, :]
    c = u[rank:, :]
    return [b @ s @ v, c @ np.linalg.inv(b)]


def get_model_singular_values(model: Model):
    """old_versions/svd_classes.py:220-232, per gate block as named (the original SVDs the
    merged matrix for every gate -- a defect)."""
    L = len(model.cells)
    H = model.cells[0].units
    out = np.zeros((L, 2, 4, H))
    for i, cell in enumerate(model.cells):
        W, U, _ = cell.get_weights()
        for j, m in enumerate((W, U)):
            for k in range(4):
                s = np.linalg.svd(m[:, H * k:H * (k + 1)], compute_uv=False)
                out[i, j, k, :s.size] = s
    return out


def set_model_matrix_rank(model: Model, index, rank):
    """old_versions/svd_classes.py:210-217; index = (cell, W|U, gate)."""
    cell = model.cells[index[0]]
    H = cell.units
    m = (cell.W, cell.U)[index[1]]
    m[:, H * index[2]:H * (index[2] + 1)] = reduce_matrix_rank(m[:, H * index[2]:H * (index[2] + 1)], rank)
    return model


def greedy_sigma_sweep(model: Model, X, y, reductions, skip_first_layer_W=False):
    """old_versions/svd_acceleration.py:61-88 (== LSTM_wrapper.iterate_reduce_model, svd_classes.py:139-182)
    without a scaler: global ascending sigma order, per iteration predict -> RMSE (sqrt of the mean over
    sequences of the per-sequence MSE), then lower the rank of the gate block owning the next-smallest sigma."""
    sv = get_model_singular_values(model)
    units = sv.shape[-1]
    idx = np.squeeze(np.dstack(np.unravel_index(np.argsort(sv.ravel(), kind="stable"), sv.shape)))
    if skip_first_layer_W:
        idx = idx[~np.logical_and(idx[:, 0] == 0, idx[:, 1] == 0)]
    ranks = np.full(sv.shape[:3], units)
    rmse_out, weights = np.zeros(reductions), np.zeros(reductions)
    running = 0
    yt = np.asarray(y, np.float64).reshape(np.shape(y)[0], -1)
    for i in range(reductions):
        yp = np.asarray(model.predict(X), np.float64).reshape(yt.shape)
        rmse_out[i] = np.sqrt(np.mean(np.mean((yt - yp) ** 2, axis=1)))
        weights[i] = running
        ci, mi, gi = (int(v) for v in idx[i][:3])
        ranks[ci, mi, gi] -= 1
        set_model_matrix_rank(model, (ci, mi, gi), int(ranks[ci, mi, gi]))
        running += 2 * units - 2 * int(ranks[ci, mi, gi]) - 1
    return rmse_out, weights


# --------------------------------------------------------------------------------------
# fixture I/O: the shipped code/model_weights layout (transposed)
# --------------------------------------------------------------------------------------
def load_model_weights_csv(path, layer_names=None, transposed=True, dtype=np.float32):
    """Returns ([(W (D,4H), U (H,4H), b (4H,)), ...], (dense_kernel (H,1), dense_bias (1,))).
    The shipped CSVs store W{g} as (units, input_dim), i.e. transposed relative to Keras."""
    if layer_names is None:
        layer_names = sorted(d for d in os.listdir(path) if d.startswith("lstm"))
    layers = []
    for name in layer_names:
        d = os.path.join(path, name)

        def rd(fn):
            return np.atleast_2d(np.loadtxt(os.path.join(d, fn), delimiter=",", dtype=np.float64))
        Ws = [rd("W%s.csv" % g) for g in GATES]
        Us = [rd("U%s.csv" % g) for g in GATES]
        bs = [np.loadtxt(os.path.join(d, "b%s.csv" % g), delimiter=",", dtype=np.float64).ravel() for g in GATES]
        if transposed:
            Ws = [w.T for w in Ws]
            Us = [u.T for u in Us]
        layers.append((np.concatenate(Ws, 1).astype(dtype), np.concatenate(Us, 1).astype(dtype),
                       np.concatenate(bs).astype(dtype)))
    dk = np.loadtxt(os.path.join(path, "dense_top", "weights.csv"), delimiter=",", dtype=np.float64).reshape(-1, 1)
    db = np.loadtxt(os.path.join(path, "dense_top", "bias.csv"), delimiter=",", dtype=np.float64).reshape(1)
    return layers, (dk.astype(dtype), db.astype(dtype))


def model_from_weights(layers, dense, dtype=np.float64, return_sequences=True):
    cells = [FullCell(U.shape[0], W, U, b, dtype=dtype) for (W, U, b) in layers]
    return Model(cells, dense, return_sequences=return_sequences, input_dim=layers[0][0].shape[0])


# --------------------------------------------------------------------------------------
# synthetic weights of SURVEY §8(d): Glorot W, per-gate orthogonal U, forget bias 1
# --------------------------------------------------------------------------------------
def synthetic_layers(D, H, L, seed=0, dtype=np.float32, n_out=1):
    rng = np.random.default_rng(seed)
    layers = []
    d_in = D
    for _ in range(L):
        lim = math.sqrt(6.0 / (d_in + 4 * H))
        W = rng.uniform(-lim, lim, size=(d_in, 4 * H))
        Us = []
        for _g in range(4):
            q, r = np.linalg.qr(rng.standard_normal((H, H)))
            q = q * np.sign(np.diag(r))
            Us.append(q)
        U = np.concatenate(Us, 1)
        b = np.zeros(4 * H)
        b[H:2 * H] = 1.0
        layers.append((W.astype(dtype), U.astype(dtype), b.astype(dtype)))
        d_in = H
    lim = math.sqrt(6.0 / (H + n_out))
    dense = (rng.uniform(-lim, lim, size=(H, n_out)).astype(dtype), np.zeros(n_out, dtype))
    return layers, dense
