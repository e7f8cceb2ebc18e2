#!/usr/bin/env python
"""BASELINE configs[3] (C4): rank x sequence sweep sharded over the GPUs of one box.

    python scripts/sweep_bench.py [--sequences 65536] [--seq-len 200] [--ranks 1:256] [--engine tc]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/sweep_bench.py ...

Synthetic 2-layer H=256 model (the C3 model), T=200 (the reference's window, svd_acceleration_v3.py:113).  Every
process (one per GPU) evaluates EVERY rank-truncated model on its contiguous shard of the sequences, reduces the
squared error against the full model's output on device (K4), and only then exchanges: one all_gather of the
last-step predictions and of the float64 SSE partials (sweep.py).  Strong scaling: the total work is fixed.
Timed region: full-model targets + all ranks + the exchange, inputs resident in HBM; max over ranks."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import svdlstm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sequences", type=int, default=65536)
    ap.add_argument("--seq-len", type=int, default=200)
    ap.add_argument("--ranks", default="1:256")
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--engine", default="tc")
    ap.add_argument("--form", default="singular")
    ap.add_argument("--target-engine", default=None, help="engine for the full-model targets (default: FP32 engines)")
    a = ap.parse_args()
    lo_r, hi_r = (int(v) for v in a.ranks.split(":"))
    ranks = list(range(lo_r, hi_r + 1))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    layers, dense = svdlstm.synthetic_layers(16, a.hidden, 2, seed=0)
    full = svdlstm.full_model_from_weights(layers, dense, return_sequences=True)
    t0 = time.perf_counter()
    smodel, models = svdlstm.build_rank_models(full, ranks, form=a.form)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    N, T = a.sequences, a.seq_len
    lo, hi = svdlstm.shard_bounds(N, world, rank)

    class Shard:                      # every process generates only its own sequences (seeded per sequence block)
        n_sequences = N

        def __call__(self, l, h):          # blocks of 1024 sequences seeded by block index: the data do not depend on the sharding
            parts = []
            for blk in range(l // 1024, (h + 1023) // 1024):
                g = torch.Generator(device=dev).manual_seed(1000 + blk)
                xb = torch.randn(1024, T, 16, generator=g, device=dev)
                parts.append(xb[max(l - blk * 1024, 0):min(h - blk * 1024, 1024)])
            return torch.cat(parts, 0)

    X = Shard()
    x_loc = X(lo, hi)
    holder = {"x": x_loc}
    Xc = type("Resident", (), {"n_sequences": N, "__call__": lambda self, l, h: holder["x"]})()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up: two ranks on a slice (kernel images, workspaces, NCCL communicator)
    svdlstm.rank_sweep(full, Xc, ranks[:2], models=models[:2], engine=a.engine, last_step_only=True, sse_over="all", target_engine=a.target_engine)
    barrier()
    l0 = svdlstm.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = svdlstm.rank_sweep(full, Xc, ranks, models=models, engine=a.engine, last_step_only=True, sse_over="all", target_engine=a.target_engine)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms[0])
    if rank == 0:
        items = len(ranks) * N
        sel = [r for r in (1, 2, 4, 8, 16, 32, 64, 128, 192, 256) if lo_r <= r <= hi_r]
        line = {"metric": "rank x sequence sweep (all ranks x sequences, RMSE vs full model)", "value": items * T / (ms * 1e-3),
                "unit": "sequence-timesteps/s (summed over ranks)", "n_gpus": world, "seconds": ms * 1e-3, "scaling": "strong",
                "items_rank_x_sequence": items, "sequences": N, "seq_len": T, "ranks": [lo_r, hi_r], "engine": a.engine, "target_engine": a.target_engine or "fp32", "form": a.form,
                "gathered_bytes": int(res["preds"].numel() * 4) if res["preds"] is not None else 0,
                "model_build_s": round(build_s, 3), "gpu_launches": svdlstm.launches() - l0,
                "rmse_vs_full": {str(r): float(res["rmse"][r - lo_r]) for r in sel}, "data": "synthetic"}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
