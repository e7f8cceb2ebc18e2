"""Time of the FP32 (CUDA-core, 1e-5-parity) engine on the headline configuration C3 and on the C4 target shape, per
sequences-per-CTA choice (SVDLSTM_GEN_BT; unset = the library's own choice)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import svdlstm
import bench

class A: hidden = 256; layers = 2; rank = 128
_, _, model = bench.build_workload(A, svdlstm)
full = model._full_parent
for bt in (None, 16, 8, 4):
    if bt is None: os.environ.pop("SVDLSTM_GEN_BT", None)
    else: os.environ["SVDLSTM_GEN_BT"] = str(bt)
    for B, T, it in ((4096, 256, 3), (512, 256, 3)):
        x = torch.randn(B, T, 16, device="cuda")
        ms = bench.timed(torch, lambda: model(x, engine="fp32"), it, warm=1)
        print("BT %s  fp32 engine rank 128  B %5d T %4d  %.2f ms  %.2f M seq-steps/s" % (bt, B, T, ms, B * T / ms / 1e3), flush=True)
    x = torch.randn(4096, 200, 16, device="cuda")
    ms = bench.timed(torch, lambda: full(x, engine="fp32"), 2, warm=1)
    print("BT %s  fp32 engine full model  B 4096 T 200  %.2f ms  %.2f M seq-steps/s" % (bt, ms, 4096 * 200 / ms / 1e3), flush=True)
