"""Driver metrics of the hot path (code/svd_acceleration_v3.py:90-100, 160-170, 187-204).

``signaltonoise`` keeps the reference signature.  RMSE has two flavours: ``reference_rmse`` keeps
the reference's divisor quirk (sum over all of y, divided by len(y_test), :187-190), ``rmse`` is the
true one.  The per-rank squared error of a sweep is reduced on device (K4, ``sweep_sse``).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _cabi as C


def signaltonoise(signal, noisy_signal, invert=False, dB=True):
    """svd_acceleration_v3.py:90-100: SNR = (A_signal/A_noise)_rms^2, in dB by default."""
    signal = np.asarray(signal, np.float64)
    noisy_signal = np.asarray(noisy_signal, np.float64)
    noise = signal - noisy_signal
    a_sig = math.sqrt(np.mean(np.square(signal)))
    a_noise = math.sqrt(np.mean(np.square(noise)))
    if not invert:
        snr = (a_sig / a_noise) ** 2
    else:
        snr = (a_noise / a_sig) ** 2
    if not dB:
        return snr
    return 10 * math.log(snr, 10)


def sweep_sse(pred, target) -> torch.Tensor:
    """Device K4: pred (R, n) [or (n,)], target (n) -> float64 CUDA tensor (R,) of sum (pred-target)^2."""
    p = C.dev_tensor(pred)
    t = C.dev_tensor(target).reshape(-1)
    if p.dim() == 1:
        p = p.unsqueeze(0)
    p = p.reshape(p.shape[0], -1)
    if p.shape[1] != t.numel():
        raise ValueError("sweep_sse: pred has %d values per rank, target has %d" % (p.shape[1], t.numel()))
    out = torch.empty(p.shape[0], dtype=torch.float64, device=p.device)
    C.check(C.lib().svdlstm_sweep_sse(C.ptr(p), C.ptr(t), int(p.shape[0]), int(t.numel()), C.ptr(out), C.cur_stream()))
    C.add_launches(2)
    return out


def rmse(y_true, y_pred) -> float:
    sse = float(sweep_sse(np.asarray(y_pred, np.float32).reshape(1, -1), np.asarray(y_true, np.float32).reshape(-1))[0])
    return math.sqrt(sse / np.size(y_true))


def reference_rmse(y_true, y_pred, n_test) -> float:
    """svd_acceleration_v3.py:188: sqrt(sum over ALL of y / len(y_test))."""
    sse = float(sweep_sse(np.asarray(y_pred, np.float32).reshape(1, -1), np.asarray(y_true, np.float32).reshape(-1))[0])
    return math.sqrt(sse / n_test)


def count_weights(model) -> int:
    """svd_acceleration_v3.py:160-166: sum of sizes of every array of model.get_weights()."""
    return int(sum(np.size(w) for w in model.get_weights()))


def weight_reduction_percent(full_model, reduced_model) -> float:
    """svd_acceleration_v3.py:170."""
    return 100 - count_weights(reduced_model) / count_weights(full_model) * 100


# closed forms of slides 8-9
def full_weight_count(D, H):
    return 4 * (D * H + H * H + H)


def reduced_split_weight_count(D, H, rw, ru):
    return 4 * (rw * (D + H - rw) + ru * (2 * H - ru)) + 4 * H


def reduced_merged_weight_count(D, H, rw, ru):
    return rw * (D + 4 * H - rw) + ru * (5 * H - ru) + 4 * H
