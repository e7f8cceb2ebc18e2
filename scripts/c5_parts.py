#!/usr/bin/env python
"""BASELINE configs[4] (C5) on ONE GPU = one rank's shard of the 8-GPU job: factorisation of the 3-layer H=1024 model
(K2: six matrices up to 1024 x 4096), rank-128 truncation, ONE fused penalty launch (K3) over all 12 factor matrices +
6 sigma vectors (device-resident, as in a training loop; checked against the oracle on the largest item), and the
recurrent forward of the shard (1024 of the 8192 sequences x T=4096) on the tensor-core engine (units = 1024 path:
16 epilogue warps, streamed weights), checked against the FP32 engine on a sub-sample.  argv: [T] [B]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import svdlstm  # noqa: E402
import svdlstm_oracle as O  # noqa: E402  (checker only)


def main():
    H, L, r = 1024, 3, 128
    layers, dense = svdlstm.synthetic_layers(16, H, L, seed=0)
    full = svdlstm.full_model_from_weights(layers, dense, return_sequences=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sm = svdlstm.make_LSTM_singular_model(full, hoyer=0.01, orthogonal=0.1, merged_kernel=True, return_sequences=True)
    torch.cuda.synchronize()
    t_svd = time.perf_counter() - t0
    tm = svdlstm.truncate_singular_model(sm, r)
    spec = []
    for layer in tm.layers[:-1]:
        w = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in layer.get_weights()]   # factors live on the device in a training loop
        spec += [(w[0], False, False), (w[1], False, False)] + [(w[i], True, False) for i in (2, 3, 4, 5)]
    svdlstm.evaluate_penalties(spec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = svdlstm.launches()
    e0.record()
    raw = svdlstm.evaluate_penalties(spec)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    nbytes = sum(a.numel() * 4 for a, _, _ in spec)
    big = max(range(len(spec)), key=lambda i: spec[i][0].numel())
    ref = O.penalty_raw_sums(spec[big][0].cpu().numpy(), mode="rows")
    # L1 / L2 sums relative; Gram sums of these (orthonormal-row) factors are ~0: absolute per pair / per entry
    n_rows = spec[big][0].shape[0]
    rel = max(abs(raw[big][0] - ref[0]) / abs(ref[0]), abs(raw[big][1] - ref[1]) / abs(ref[1]),
              abs(raw[big][2] - ref[2]) / (n_rows * (n_rows - 1)), abs(raw[big][3] - ref[3]) / (n_rows * n_rows))
    out = {"config": "C5 shard of one GPU: L=3 H=1024 rank 128", "svd_factorisation_s": round(t_svd, 3),
           "penalty_items": len(spec), "penalty_launches": svdlstm.launches() - l0, "penalty_ms": round(ms, 3),
           "penalty_bytes_read_once": nbytes, "penalty_GBps": round(nbytes / ms / 1e6, 1),
           "largest_item_shape": list(spec[big][0].shape), "max_err_vs_oracle (L1, L2 relative; Gram sums absolute per entry)": rel}
    # ---- the recurrent forward of this GPU's shard on the tensor-core engine
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    x = torch.randn(B, T, 16, generator=torch.Generator().manual_seed(0)).cuda()
    y = tm(x, engine="tc")           # warm-up: packs the weight streams, allocates the hand-off images
    torch.cuda.synchronize()
    e0.record()
    y = tm(x, engine="tc")
    e1.record()
    torch.cuda.synchronize()
    fms = e0.elapsed_time(e1)
    macs, d = 0, 16
    for _ in range(L):
        macs += min(r, d) * (d + 4 * H) + min(r, H) * 5 * H
        d = H
    sub = x[:32, :24].contiguous()
    err = (tm(sub, engine="tc") - tm(sub, engine="general")).abs().max().item()
    scale = tm(sub, engine="general").abs().max().item()
    out.update({"forward_B": B, "forward_T": T, "forward_ms": round(fms, 2), "forward_Mseqsteps_per_s": round(B * T / fms / 1e3, 2),
                "forward_algorithmic_TFLOPs": round(2 * macs * B * T / (fms * 1e-3) / 1e12, 1),
                "forward_frac_of_1384": round(2 * macs * B * T / (fms * 1e-3) / 1e12 / 1384, 4),
                "forward_max_abs_tc_minus_fp32_subsample": err, "forward_subsample_output_max": scale,
                "forward_finite": bool(torch.isfinite(y).all().item())})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
