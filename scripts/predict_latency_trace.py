"""Per-call latency of blocking model.predict(pinned host x) at the headline shape (sorted list of 20 calls), for plain pinned and
write-combined staging."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, svdlstm, bench
class A: hidden=256; layers=2; rank=128
_,_,model = bench.build_workload(A, svdlstm)
src = torch.randn(4096,1024,16)
for wc in (False, True):
    x = svdlstm.pinned_empty((4096,1024,16), write_combined=wc); x.copy_(src)
    model.predict(x); model.predict(x); torch.cuda.synchronize()
    ts=[]
    for _ in range(20):
        t0=time.perf_counter(); y=model.predict(x); ts.append((time.perf_counter()-t0)*1e3)
    print("write_combined=%s" % wc, " ".join("%.2f" % t for t in sorted(ts)), flush=True)
