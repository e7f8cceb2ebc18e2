"""Bring-up check of the tcgen05 engine against the FP32 general engine (same weights, same inputs)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import svdlstm  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = int(sys.argv[2]) if len(sys.argv) > 2 else 2
r = int(sys.argv[3]) if len(sys.argv) > 3 else 16
B = int(sys.argv[4]) if len(sys.argv) > 4 else 64
T = int(sys.argv[5]) if len(sys.argv) > 5 else 8
form = sys.argv[6] if len(sys.argv) > 6 else "3F"
layers, dense = svdlstm.synthetic_layers(16, H, L, seed=0)
full = svdlstm.full_model_from_weights(layers, dense)
sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=True, return_sequences=True)
model = svdlstm.truncate_singular_model(sm, r) if form == "3F" else svdlstm.make_LSTM_reduced_model(sm, rank=r)
x = torch.randn(B, T, 16, generator=torch.Generator().manual_seed(0)).cuda()
y32 = model(x, engine="general")
torch.cuda.synchronize()
ytc = model(x, engine="tc")
torch.cuda.synchronize()
d = (ytc - y32).abs()
print("H=%d L=%d r=%d B=%d T=%d %s" % (H, L, r, B, T, form))
print("fp32 |y| mean %.4f max %.4f" % (y32.abs().mean().item(), y32.abs().max().item()))
print("tc-vs-fp32: max abs %.4e  mean abs %.4e  rmse %.4e" % (d.max().item(), d.mean().item(), (d * d).mean().sqrt().item()))
print("per-step max err:", [round(v, 5) for v in d.amax(dim=(0, 2)).tolist()])
print("nan:", torch.isnan(ytc).any().item())
