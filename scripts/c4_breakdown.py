"""Where the time of the C4 sweep (ranks 1..256 x 65 536 sequences x T=200) goes on one GPU: FP32 targets, the 256 truncated
models (device time and host wall time of the loop: a gap = the host is the bottleneck), SSE reductions."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import svdlstm

R, N, T = 256, 65536, 200
layers, dense = svdlstm.synthetic_layers(16, 256, 2, seed=0)
full = svdlstm.full_model_from_weights(layers, dense, return_sequences=True)
t0 = time.perf_counter()
_, models = svdlstm.build_rank_models(full, list(range(1, R + 1)), form="singular")
torch.cuda.synchronize()
print("build %.3f s" % (time.perf_counter() - t0), flush=True)
x = torch.randn(N, T, 16, device="cuda")
ev = lambda: torch.cuda.Event(enable_timing=True)
for m in models[:2]:
    m(x)
full(x[:4096], engine="fp32")
torch.cuda.synchronize()
e0, e1 = ev(), ev()
e0.record(); tgt = full(x, engine="fp32"); e1.record(); torch.cuda.synchronize()
print("fp32 targets: %.3f s" % (e0.elapsed_time(e1) * 1e-3), flush=True)
for label, ms_ in (("ranks 1..64", models[:64]), ("ranks 65..128", models[64:128]), ("ranks 129..256", models[128:])):
    e0, e1 = ev(), ev()
    w0 = time.perf_counter()
    e0.record()
    for m in ms_:
        y = m(x)
    w_issue = time.perf_counter() - w0
    e1.record(); torch.cuda.synchronize()
    print("%-15s device %.3f s   host issue %.3f s   (%d models)" % (label, e0.elapsed_time(e1) * 1e-3, w_issue, len(ms_)), flush=True)
# second pass over the same (now packed) models: no weight packing on the way
e0, e1 = ev(), ev()
e0.record()
for m in models:
    y = m(x)
e1.record(); torch.cuda.synchronize()
print("all 256 again (weights already packed): %.3f s" % (e0.elapsed_time(e1) * 1e-3))
