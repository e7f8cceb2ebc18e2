"""Times the tensor-core engine for a list of ranks.  argv: ranks B T H L (defaults: the C3 workload shape)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import svdlstm  # noqa: E402

ranks = [int(r) for r in sys.argv[1].split(",")] if len(sys.argv) > 1 else [8, 16, 32]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
H = int(sys.argv[4]) if len(sys.argv) > 4 else 256
L = int(sys.argv[5]) if len(sys.argv) > 5 else 2
layers, dense = svdlstm.synthetic_layers(16, H, L, seed=0)
full = svdlstm.full_model_from_weights(layers, dense)
sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=True, return_sequences=True)
x = torch.randn(B, T, 16, generator=torch.Generator().manual_seed(0)).cuda()
for r in ranks:
    model = svdlstm.truncate_singular_model(sm, r)
    for _ in range(2):
        y = model(x, engine="tc")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        y = model(x, engine="tc")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    macs = 0
    d = 16
    for _ in range(L):
        macs += min(r, d) * (d + 4 * H) + min(r, H) * 5 * H
        d = H
    tf = 2 * macs * B * T / (ms * 1e-3) / 1e12
    ysub = model(x[:64, :64], engine="general")
    err = (model(x[:64, :64], engine="tc") - ysub).abs().max().item()
    print("rank %3d: %.3f ms/step  %.2f M seq-steps/s  %.1f TFLOP/s (algorithmic)  max|tc-fp32| %.2e" % (r, ms, B * T / ms / 1e3, tf, err))
