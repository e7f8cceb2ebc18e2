#!/usr/bin/env python
"""BASELINE configs[4] (C5) on the GPUs of one box: 3-layer SVD-LSTM H=1024, seq_len=4096, batch=8192, rank 128 + the fused
penalty evaluation.  Pure data parallelism over the batch: every process (one per GPU) runs batch/world sequences through
the tensor-core engine; the only exchange is the final all_gather of the outputs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/c5_bench.py [--seq-len 4096] [--batch 8192]

Timed (CUDA events, max over ranks, barrier + synchronize on both sides): the forward of the shard, the gather, the penalties."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import svdlstm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seq-len", type=int, default=4096)
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--rank", type=int, default=128)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    H, L, r, T = 1024, 3, a.rank, a.seq_len
    lo, hi = svdlstm.shard_bounds(a.batch, world, rank)
    layers, dense = svdlstm.synthetic_layers(16, H, L, seed=0)
    full = svdlstm.full_model_from_weights(layers, dense, return_sequences=True)
    sm = svdlstm.make_LSTM_singular_model(full, hoyer=0.01, orthogonal=0.1, merged_kernel=True, return_sequences=True)
    tm = svdlstm.truncate_singular_model(sm, r)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    x = torch.randn(hi - lo, T, 16, generator=g, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return out, float(ms[0])

    tm(x[:64], engine="tc")
    tm(x, engine="tc")                      # warm-up at full size (workspaces)
    y, fwd_ms = timed(lambda: tm(x, engine="tc"))

    def gather():
        if dist is None:
            return y
        bufs = [torch.empty_like(y) for _ in range(world)]
        dist.all_gather(bufs, y)
        return torch.cat(bufs, 0)

    gather()
    y_all, gat_ms = timed(gather)
    spec = []
    for layer in tm.layers[:-1]:
        w = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in layer.get_weights()]
        spec += [(w[0], False, False), (w[1], False, False)] + [(w[i], True, False) for i in (2, 3, 4, 5)]
    svdlstm.evaluate_penalties(spec)
    _, pen_ms = timed(lambda: svdlstm.evaluate_penalties(spec))
    if rank == 0:
        macs, d = 0, 16
        for _ in range(L):
            macs += min(r, d) * (d + 4 * H) + min(r, H) * 5 * H
            d = H
        tf = 2 * macs * a.batch * T / (fwd_ms * 1e-3) / 1e12
        print(json.dumps({"config": "C5: L=3 H=1024 T=%d batch=%d rank=%d" % (T, a.batch, r), "n_gpus": world, "sequences_per_gpu": hi - lo,
                          "forward_ms": round(fwd_ms, 2), "sequence_steps_per_s": round(a.batch * T / (fwd_ms * 1e-3)),
                          "algorithmic_TFLOPs_all_gpus": round(tf, 1), "frac_of_peak_per_gpu": round(tf / world / 1384, 4),
                          "gather_ms": round(gat_ms, 2), "gathered_bytes": int(y_all.numel() * 4), "penalties_ms": round(pen_ms, 3),
                          "output_finite": bool(torch.isfinite(y_all).all().item()), "data": "synthetic"}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
