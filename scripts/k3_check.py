"""K3 check: tensor-core Gram tiles vs the float64 CUDA-core tiles vs the numpy oracle, item by item."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import svdlstm  # noqa: E402
import svdlstm_oracle as O  # noqa: E402  (checker only)

rng = np.random.default_rng(20)
q = np.linalg.qr(rng.standard_normal((256, 256)))[0].astype(np.float32)
items = [(rng.standard_normal((128, 4096)) / 64.0).astype(np.float32), (rng.standard_normal((1024, 128)) / 11.0).astype(np.float32),
         rng.standard_normal((300, 700)).astype(np.float32), q, q[:100, :].copy(), rng.standard_normal((64, 64)).astype(np.float32),
         rng.standard_normal((129, 65)).astype(np.float32), rng.standard_normal((130, 257)).astype(np.float32)]
spec = [(torch.from_numpy(a).cuda(), True, False) for a in items]
raw = svdlstm.evaluate_penalties(spec)
os.environ["SVDLSTM_K3_NO_TC"] = "1"
raw64 = svdlstm.evaluate_penalties(spec)
del os.environ["SVDLSTM_K3_NO_TC"]
for a, g, g64 in zip(items, raw, raw64):
    ref = O.penalty_raw_sums(a, mode="rows")
    print(a.shape, "off: tc %.9g f64 %.9g ref %.9g | fro: tc %.9g f64 %.9g ref %.9g" % (g[2], g64[2], ref[2], g[3], g64[3], ref[3]))
for label, env in (("tensor cores", None), ("float64 CUDA cores", "1")):
    if env:
        os.environ["SVDLSTM_K3_NO_TC"] = env
    svdlstm.evaluate_penalties(spec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(10):
        svdlstm.evaluate_penalties(spec)
    e1.record()
    torch.cuda.synchronize()
    print("%s: %.3f ms per call (events), %.3f ms wall" % (label, e0.elapsed_time(e1) / 10, (time.perf_counter() - t0) * 100))
    os.environ.pop("SVDLSTM_K3_NO_TC", None)
