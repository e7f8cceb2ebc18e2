#!/usr/bin/env python
"""BASELINE configs[4] (C5) pieces that run today at full size on one GPU: factorisation of the 3-layer H=1024 model
(K2: six matrices up to 1024 x 4096), rank-128 truncation, and ONE fused penalty launch (K3) over all 12 factor
matrices + 6 sigma vectors, checked against the oracle on the largest item.  The recurrent forward at H=1024 runs on
the FP32 general engine only (the tensor-core engine holds c in registers: units <= 512) -- see DESIGN.md section 7."""
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
        w = layer.get_weights()
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
    nbytes = sum(np.asarray(a).size * 4 for a, _, _ in spec)
    big = max(range(len(spec)), key=lambda i: np.asarray(spec[i][0]).size)
    ref = O.penalty_raw_sums(np.asarray(spec[big][0]), mode="rows")
    rel = max(abs(raw[big][k] - ref[k]) / (abs(ref[k]) + 1e-12) for k in range(4))
    print(json.dumps({"config": "C5 pieces: L=3 H=1024 rank 128", "svd_factorisation_s": round(t_svd, 3),
                      "penalty_items": len(spec), "penalty_launches": svdlstm.launches() - l0, "penalty_ms": round(ms, 3),
                      "penalty_bytes_read_once": nbytes, "penalty_GBps": round(nbytes / ms / 1e6, 1),
                      "largest_item_shape": list(np.asarray(spec[big][0]).shape), "max_rel_err_vs_oracle": rel}))


if __name__ == "__main__":
    main()
