"""Explicit-rank API of the original sweep (code/old_versions/svd_classes.py:9-21, 210-232), on device.

    reduce_matrix_rank(a, rank)        :9-12    A_r = (U*s_r) V
    reduce_two_step(a, rank)           :14-21   column-vector two-step [b s v ; c b^-1]
    get_model_singular_values(model)   :220-232 (layers, 2, 4, units) table of per-gate sigma
    set_model_matrix_rank(model, index, rank)   :210-217  re-SVD one gate block, zero trailing sigma
    LSTM_wrapper(model, scaler).iterate_reduce_model(X, y, ...)   :139-182 / old_versions/svd_acceleration.py:61-88
                                       greedy "drop the globally smallest sigma, predict, RMSE" sweep
All SVDs run through K2 (svd_batched); the reconstruction (U * s_r) V runs through K5 (scaled_matmul).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _cabi as C
from .models import reduce_factors, svd_batched


def reduce_matrix_rank(a, rank):
    A = C.dev_tensor(a)
    u, s, v = svd_batched(A)
    r = max(0, min(int(rank), int(s.numel())))
    return C.scaled_matmul(u, v, scale=s.reshape(-1), k=r).cpu().numpy()      # (U * s_r) V on device (K5)


def reduce_two_step(a, rank):
    """Returns [M1 (r,m), M2 (n-r,r)] with y[:r] = M1 x, y[r:] = M2 y[:r] for A (n,m) (column vectors).
    Transposed twin of make_LSTM_reduced_model's (B, C): M1 = B^T, M2 = C^T of A^T."""
    A = C.dev_tensor(a)
    rank = int(rank)
    u, s, v = svd_batched(A.t().contiguous())        # A^T = u s v ;  u (m,k) v (k,n)
    B, Cm = reduce_factors(u[:, :rank].contiguous(), s[:rank].contiguous(), v[:rank].contiguous())
    return [B.t().contiguous().cpu().numpy(), Cm.t().contiguous().cpu().numpy()]


def get_model_singular_values(model):
    """Per (layer, W|U, gate) singular values (old_versions/svd_classes.py:220-232).  The original SVDs
    the whole merged matrix for every gate (a defect); the per-gate block is used here."""
    lstms = model.layers[:-1]
    units = lstms[0].units
    out = np.zeros((len(lstms), 2, 4, units))
    for i, layer in enumerate(lstms):
        W, U = layer.cell.kernel.tensor, layer.cell.recurrent_kernel.tensor
        for j, m in enumerate((W, U)):
            rows = int(m.shape[0])
            blocks = m.view(rows, 4, layer.units).permute(1, 0, 2).contiguous()
            s = svd_batched(blocks, compute_uv=False).cpu().numpy()
            out[i, j, :, :s.shape[1]] = s
    return out


def set_model_matrix_rank(model, index, rank):
    """index = (cell, W|U, gate) (old_versions/svd_classes.py:210-217)."""
    layer = model.layers[index[0]]
    H = layer.units
    var = (layer.cell.kernel, layer.cell.recurrent_kernel)[index[1]]
    blk = var.tensor[:, H * index[2]:H * (index[2] + 1)]
    u, s, v = svd_batched(blk.contiguous())
    r = max(0, min(int(rank), int(s.numel())))
    C.scaled_matmul(u, v, scale=s.reshape(-1), k=r, out=blk)                   # written in place into the gate block (strided view)
    layer.cell.rebind()
    return model


def sorted_sigma_indices(singular_values, skip_first_layer_W=False):
    """Global ascending order of every (cell, W|U, gate, k) singular value
    (old_versions/svd_acceleration.py:61-64); optionally without the first layer's W (:66-68)."""
    sv = np.asarray(singular_values)
    idx = np.squeeze(np.dstack(np.unravel_index(np.argsort(sv.ravel(), kind="stable"), sv.shape)))
    if skip_first_layer_W:
        idx = idx[~np.logical_and(idx[:, 0] == 0, idx[:, 1] == 0)]
    return idx


class LSTM_wrapper:
    """old_versions/svd_classes.py:139-182: wraps a full model with the greedy one-sigma-at-a-time reduction.
    Every iteration re-factors ONE gate block on device (K2) and re-runs the persistent forward kernel; the
    squared error is reduced on device (K4).  `scaler` needs `inverse_transform` (sklearn-like) or is None."""

    def __init__(self, model, scaler=None):
        self.scaler = scaler
        self.model = model
        lstms = model.layers[:-1]
        units = lstms[0].units
        self.model_ranks = np.full((len(lstms), 2, 4), units)     # :144 (hard-coded 30 upstream: "change to model dimensions")
        self.model_singular_values = get_model_singular_values(model)
        self.rmse = None
        self.weights_eliminated = None

    def _inv(self, a):
        a = np.asarray(a, np.float64)
        a = a.reshape(a.shape[0], -1)
        if self.scaler is None:
            return a
        return np.asarray(self.scaler.inverse_transform(a), np.float64)

    def evaluate(self, X, y):
        """RMSE as the reference computes it (:162-164): sqrt of the mean over sequences of per-sequence MSE."""
        from .metrics import sweep_sse
        y_true = self._inv(y)
        pred = self.model.predict(X)
        y_pred = self._inv(pred)
        if self.scaler is None:      # no host transform in between: reduce on device
            sse = float(sweep_sse(np.ascontiguousarray(y_pred, np.float32).reshape(1, -1),
                                  np.ascontiguousarray(y_true, np.float32).reshape(-1)).cpu().numpy()[0])
            return float(np.sqrt(sse / y_true.size))
        return float(np.sqrt(np.mean(np.mean((y_true - y_pred) ** 2, axis=1))))

    def iterate_reduce_model(self, X, y, threshold=None, reductions=None, evaluate_every=1, heuristic="absolute",
                             skip_first_layer_W=False):
        if heuristic != "absolute":
            raise ValueError("only the 'absolute' heuristic exists upstream")
        units = self.model_singular_values.shape[-1]
        order = sorted_sigma_indices(self.model_singular_values, skip_first_layer_W)
        iterations = order.shape[0] - 20                      # :158
        if reductions is None:
            reductions = iterations
        reductions = min(int(reductions), order.shape[0])
        rmse = np.zeros((reductions + evaluate_every - 1) // evaluate_every)
        weights_eliminated = np.zeros(reductions)
        running = 0
        done = 0
        for i in range(reductions):
            if i % evaluate_every == 0:
                rmse[i // evaluate_every] = self.evaluate(X, y)
                weights_eliminated[i] = running
                if threshold is not None and rmse[i // evaluate_every] > threshold:
                    break
            ci, mi, gi = (int(v) for v in order[i][:3])
            rank = int(self.model_ranks[ci, mi, gi]) - 1
            self.model_ranks[ci, mi, gi] = rank
            self.model = set_model_matrix_rank(self.model, (ci, mi, gi), rank)
            running += 2 * units - 2 * rank - 1               # :176
            done = i + 1
        self.rmse, self.weights_eliminated = rmse, weights_eliminated
        return rmse, weights_eliminated, done
