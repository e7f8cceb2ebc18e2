"""Explicit-rank API of the original sweep (code/old_versions/svd_classes.py:9-21, 210-232), on device.

    reduce_matrix_rank(a, rank)        :9-12    A_r = (U*s_r) V
    reduce_two_step(a, rank)           :14-21   column-vector two-step [b s v ; c b^-1]
    get_model_singular_values(model)   :220-232 (layers, 2, 4, units) table of per-gate sigma
    set_model_matrix_rank(model, index, rank)   :210-217  re-SVD one gate block, zero trailing sigma
All SVDs run through K2 (svd_batched); the tiny dense products use library GEMMs (plumbing).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _cabi as C
from .models import reduce_factors, svd_batched


def reduce_matrix_rank(a, rank):
    A = C.dev_tensor(a)
    u, s, v = svd_batched(A)
    s = s.clone()
    s[int(rank):] = 0
    return ((u * s) @ v).cpu().numpy()


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
    s = s.clone()
    s[int(rank):] = 0
    blk.copy_((u * s) @ v)
    layer.cell.rebind()
    return model
