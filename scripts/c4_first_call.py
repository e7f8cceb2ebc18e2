"""Cost of a model's FIRST forward (weight packing, allocations) against its second, synchronised on both sides."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import svdlstm

N, T = 65536, 200
layers, dense = svdlstm.synthetic_layers(16, 256, 2, seed=0)
full = svdlstm.full_model_from_weights(layers, dense, return_sequences=True)
ranks = [4, 8, 16, 24, 32, 48, 64, 96, 128, 160, 192, 224, 256] * 3
_, models = svdlstm.build_rank_models(full, ranks, form="singular")
x = torch.randn(N, T, 16, device="cuda")
models[0](x); models[-1](x)
torch.cuda.synchronize()
first, second = [], []
for m in models[1:-1]:
    for dst in (first, second):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        y = m(x)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        dst.append((t1 - t0, t2 - t0))
f, s = np.array(first) * 1e3, np.array(second) * 1e3
print("first call : host return %.2f ms (median), done %.2f ms (median), %.2f (mean)" % (np.median(f[:, 0]), np.median(f[:, 1]), f[:, 1].mean()))
print("second call: host return %.2f ms (median), done %.2f ms (median), %.2f (mean)" % (np.median(s[:, 0]), np.median(s[:, 1]), s[:, 1].mean()))
print("first - second, per model: median %.2f ms, mean %.2f ms, max %.2f ms" % (np.median(f[:, 1] - s[:, 1]), (f[:, 1] - s[:, 1]).mean(), (f[:, 1] - s[:, 1]).max()))
