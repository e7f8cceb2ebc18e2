"""Rank x sequence sweep, sharded over GPUs by independent sequences.

Generalises the reference's sweep loop -- "truncate, predict, RMSE" per rank
(code/old_versions/svd_acceleration.py:78-88; timing/RMSE cells of code/svd_acceleration_v3.py:145-194)
-- to all ranks x a batch of sequences.  Work item = (rank, sequence); items are independent, so each
rank (process, one per GPU) takes a contiguous slice of the sequences, evaluates every truncated model on
it, reduces the per-rank squared error on device (K4) and only then exchanges: one all_gather of the
predictions and one all_reduce(sum) of the float64 SSE vector.  The partition depends only on
(n_sequences, world_size), and SSE partials are summed in rank order, so results do not depend on how
many GPUs ran the sweep beyond float64 summation order of <= 8 terms.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np
import torch

from . import _cabi as C
from .metrics import sweep_sse
from .models import make_LSTM_reduced_model, make_LSTM_singular_model, truncate_singular_model


def shard_bounds(n: int, world_size: int, rank: int):
    """Contiguous, balanced partition of range(n): the first n % world_size shards get one extra item."""
    base, extra = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist
    return None


def exchange_results(preds, sse, count, N, gather_predictions=True):
    """The only communication of the sweep: all_gather of per-process SSE partials (summed in rank order
    => deterministic), of the item counts, and (optionally) of the predictions, un-padded and concatenated
    in sequence order.  Works on any backend (NCCL on GPUs; gloo in the CPU tests).  No-op without
    torch.distributed."""
    dist = _dist()
    if dist is None:
        return sse, count, (preds if gather_predictions else None)
    world = dist.get_world_size()
    parts = [torch.empty_like(sse) for _ in range(world)]
    dist.all_gather(parts, sse)
    sse = torch.stack(parts, 0).sum(0)
    cparts = [torch.empty_like(count) for _ in range(world)]
    dist.all_gather(cparts, count)
    count = torch.stack(cparts, 0).sum(0)
    all_preds = None
    if gather_predictions:
        R, n_loc, per = preds.shape
        sizes = [shard_bounds(N, world, r)[1] - shard_bounds(N, world, r)[0] for r in range(world)]
        mx = max(sizes)
        pad = torch.zeros((R, mx, per), dtype=preds.dtype, device=preds.device)
        pad[:, :n_loc] = preds
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad)
        all_preds = torch.cat([bufs[r][:, :sizes[r]] for r in range(world)], 1)
    return sse, count, all_preds


def build_rank_models(full_model, ranks: Sequence[int], form="reduced", merged_kernel=True):
    """One SVD of the full model (K2), then one truncated model per rank: 2-factor (`reduced`, K2b) or
    3-factor (`singular`)."""
    smodel = make_LSTM_singular_model(full_model, merged_kernel=merged_kernel, return_sequences=True)
    models = []
    for r in ranks:
        if form == "reduced":
            models.append(make_LSTM_reduced_model(smodel, merged_kernel=merged_kernel, rank=int(r)))
        elif form == "singular":
            models.append(truncate_singular_model(smodel, int(r)))
        else:
            raise ValueError("form must be 'reduced' or 'singular'")
    return smodel, models


def rank_sweep(full_model, X, ranks: Sequence[int], *, target=None, form="reduced", merged_kernel=True, engine=None,
               gather_predictions=True, last_step_only=False, models=None, sse_over="kept", target_engine=None):
    """Evaluate every rank-truncated model on this process's shard of X and reduce.

    X: (N, T, D) host or device array holding ALL sequences (each process slices its own shard), or, if
    ``X`` is a callable, ``X(lo, hi)`` returns the shard.  target: (N, T[,1]) reference outputs; default =
    the full model's own output (the sweep then measures truncation error, as the reference's RMSE ratio
    plot does).  Returns dict(ranks, rmse (R,), sse (R,), n, preds (R,N,T') or None, shard=(lo,hi))."""
    dist = _dist()
    world = dist.get_world_size() if dist else 1
    me = dist.get_rank() if dist else 0
    N = int(X.n_sequences) if callable(X) else int(np.shape(X)[0])
    lo, hi = shard_bounds(N, world, me)
    x_loc = X(lo, hi) if callable(X) else X[lo:hi]
    x_loc = C.dev_tensor(x_loc)
    if models is None:
        _, models = build_rank_models(full_model, ranks, form=form, merged_kernel=merged_kernel)
    if target is None:
        tgt = full_model(x_loc, engine=target_engine)   # default: the FP32 engines; "tc" compares like with like (truncation error only)
    else:
        tgt = C.dev_tensor(target(lo, hi) if callable(target) else target[lo:hi])
    tgt = tgt.reshape(tgt.shape[0], -1) if tgt.dim() > 1 else tgt.reshape(-1, 1)
    tgt_all = tgt
    if last_step_only:
        tgt = tgt[:, -1:]
    if sse_over not in ("kept", "all"):
        raise ValueError("sse_over must be 'kept' or 'all'")
    R = len(ranks)
    per = tgt.shape[1]
    per_sse = tgt_all.shape[1] if sse_over == "all" else per
    preds = torch.empty((R, hi - lo, per), dtype=torch.float32, device=x_loc.device)
    sse = torch.zeros(R, dtype=torch.float64, device=x_loc.device)
    tgt_flat = tgt_all.reshape(-1).contiguous()
    for i, m in enumerate(models):
        y = m(x_loc, engine=engine)
        y = y.reshape(y.shape[0], -1)
        preds[i] = y[:, -1:] if last_step_only else y
        if sse_over == "all" and last_step_only and hi > lo:
            sse[i:i + 1] = sweep_sse(y.reshape(1, -1), tgt_flat)
    if hi > lo and not (sse_over == "all" and last_step_only):
        sse = sweep_sse(preds.reshape(R, -1), tgt.reshape(-1))
    per = per_sse
    count = torch.tensor([float((hi - lo) * per)], dtype=torch.float64, device=x_loc.device)
    sse, count, all_preds = exchange_results(preds, sse, count, N, gather_predictions=gather_predictions)
    sse_h = sse.cpu().numpy()
    n_tot = float(count.cpu().numpy()[0])
    return {"ranks": list(ranks), "sse": sse_h, "rmse": np.sqrt(sse_h / n_tot), "n": n_tot,
            "preds": all_preds, "shard": (lo, hi), "world_size": world}
