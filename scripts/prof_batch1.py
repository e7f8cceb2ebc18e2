"""Profiling driver: shipped DROPBEAR model, batch 1, one wavefront launch (ncu target)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import svdlstm  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
form = sys.argv[2] if len(sys.argv) > 2 else "3F"
layers, dense = svdlstm.load_model_weights_npz(os.path.join(ROOT, "tests", "golden", "dropbear_weights.npz"))
full = svdlstm.full_model_from_weights(layers, dense)
sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=True, return_sequences=True)
model = {"full": full, "3F": sm, "2F": svdlstm.make_LSTM_reduced_model(sm, rank=8)}[form]
x = torch.randn(1, T, 16, generator=torch.Generator().manual_seed(0)).cuda()
for _ in range(3):
    y = model(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
y = model(x)
e1.record()
torch.cuda.synchronize()
print("engine", model.last_engine(), "us/step %.4f" % (e0.elapsed_time(e1) * 1e3 / T))
