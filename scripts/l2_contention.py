"""Is the weight stream of the tensor-core kernel throttled by L2 when all tiles run at once?  Same kernel (64-sequence tiles,
pipelined launch), same per-CTA work, 16..64 tiles per layer: if the step time grows with the number of co-running CTAs the
shared resource (L2 -> SM) is a co-bound."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SVDLSTM_TC_NS", "64")
import torch
import svdlstm
import bench

class A: hidden = 256; layers = 2
T = 256
for rank in (128, 256, 32):
    A.rank = rank
    _, _, model = bench.build_workload(A, svdlstm)
    for B in (512, 1024, 2048, 3072, 4096, 4608):
        x = torch.randn(B, T, 16, device="cuda")
        f = lambda: model(x, engine="tc")
        ms = bench.timed(torch, f, 10, warm=3)
        print("rank %3d  B %4d (%2d tiles/layer)  %.3f ms  %.0f cycles/step at 1.965 GHz" % (rank, B, (B + 63) // 64, ms, ms * 1e-3 / T * 1.965e9), flush=True)
