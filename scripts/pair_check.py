"""Bring-up check of the paired-CTA (cta_group::2) tensor-core kernel against the FP32 general engine and the single-CTA
tensor-core kernel (same weights, same inputs).  argv: H L rank B T [form]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import svdlstm  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = int(sys.argv[2]) if len(sys.argv) > 2 else 2
r = int(sys.argv[3]) if len(sys.argv) > 3 else 128
B = int(sys.argv[4]) if len(sys.argv) > 4 else 128
T = int(sys.argv[5]) if len(sys.argv) > 5 else 8
form = sys.argv[6] if len(sys.argv) > 6 else "3F"
layers, dense = svdlstm.synthetic_layers(16, H, L, seed=0)
full = svdlstm.full_model_from_weights(layers, dense)
sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=True, return_sequences=True)
model = svdlstm.truncate_singular_model(sm, r) if form == "3F" else svdlstm.make_LSTM_reduced_model(sm, rank=r)
x = torch.randn(B, T, 16, generator=torch.Generator().manual_seed(0)).cuda()
y32 = model(x, engine="general")
torch.cuda.synchronize()
os.environ.pop("SVDLSTM_TC_MODE", None)
ytc = model(x, engine="tc_bf16")
torch.cuda.synchronize()
os.environ["SVDLSTM_TC_MODE"] = "pair"
yp = model(x, engine="tc_bf16")
torch.cuda.synchronize()
yp2 = model(x, engine="tc_bf16")
torch.cuda.synchronize()
d = (yp - y32).abs()
d1 = (ytc - y32).abs()
print("H=%d L=%d r=%d B=%d T=%d %s" % (H, L, r, B, T, form))
print("fp32 |y| mean %.4f max %.4f" % (y32.abs().mean().item(), y32.abs().max().item()))
print("single-CTA tc vs fp32: max abs %.4e rmse %.4e" % (d1.max().item(), (d1 * d1).mean().sqrt().item()))
print("pair tc vs fp32:       max abs %.4e rmse %.4e   rerun identical: %s" % (d.max().item(), (d * d).mean().sqrt().item(), torch.equal(yp, yp2)))
print("per-step max err:", [round(v, 5) for v in d.amax(dim=(0, 2)).tolist()[:16]])
print("per-128-tile max err:", [round(v, 5) for v in d.view(-1, T).reshape(B, T).amax(dim=1).view(-1)[: (B // 32) * 32].view(-1, 32).amax(dim=1).tolist()[:16]])
print("nan:", torch.isnan(yp).any().item())
