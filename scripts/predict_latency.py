"""Latency of ONE blocking model.predict(pinned host x) at the headline shape, per number of upload slices (SVDLSTM_INPUT_SLICES;
0 = slices of 64, 64, 128, 256 ... steps)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, svdlstm, bench
class A: hidden=256; layers=2; rank=int(sys.argv[1]) if len(sys.argv) > 1 else 128
_,_,model = bench.build_workload(A, svdlstm)
x = svdlstm.pinned_empty((4096,1024,16)); x.copy_(torch.randn(4096,1024,16))
for ns in ("32", "8", "16", "64", "0", "32") if len(sys.argv) < 2 else ("32",):
    os.environ["SVDLSTM_INPUT_SLICES"]=ns
    model.predict(x); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(5): y=model.predict(x)
    torch.cuda.synchronize()
    dt=(time.perf_counter()-t0)/5*1e3
    # without the owned numpy copy: predict_async().result()
    t0=time.perf_counter()
    for _ in range(5): y=model.predict_async(x).result()
    torch.cuda.synchronize()
    dt2=(time.perf_counter()-t0)/5*1e3
    print("slices %s: predict %.2f ms   predict_async().result() %.2f ms" % (ns, dt, dt2), flush=True)
