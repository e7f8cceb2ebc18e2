#!/usr/bin/env python
"""bench.py -- headline benchmark of the SVD-factored LSTM hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
                    [--rank R] [--engine tc|general|auto] [--batch B] [--seq-len T] [--quick]

Workload (BASELINE.json configs[2], the config the metric "low-rank LSTM timesteps/sec (batch 4096)" is
quoted on): synthetic 2-layer SVD-LSTM, D=16, H=256, seq_len=1024, batch=4096 per GPU, 3-factor form
truncated to rank R (default 128), Dense(1) top, run on the tcgen05 tensor-core engine (FP16 operands,
FP32 accumulation and cell state; its RMSE delta against the FP32 parity engine is measured in the same run
and reported under "rmse_delta").  One "step" = one forward pass of the whole batch through all T
timesteps, all layers and the Dense top; "rank_sweep" repeats the measurement for ranks 8..256.  value = sequence-timesteps/s over all GPUs
(weak scaling: every GPU runs its own 4096 sequences; the path shards by independent sequences with
no data-path collective).  The batch-1 half of the metric (us/step vs rank on the shipped DROPBEAR
model, configs[1]) is measured in the same run and reported under "batch1_us_per_step".

The CPU oracle is used ONLY for the `cpu_baseline` leg and for `--impl reference` (a bounded sample of
the same workload on the host cores); it is never on the measured GPU path.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--rank", type=int, default=128)
    ap.add_argument("--engine", default="tc")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--seq-len", type=int, default=1024)
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--quick", action="store_true", help="small shapes (CI / profiling)")
    ap.add_argument("--no-batch1", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    a = ap.parse_args()
    if a.quick:
        a.batch, a.seq_len = 512, 64
    return a


def flops_per_seq_step(D, H, L, r, form="singular"):
    """Algorithmic GEMM FLOPs per sequence-timestep: 2*sum_l[r_w(D_l+4H) + r_u*5H]."""
    macs = 0
    d = D
    for _ in range(L):
        rw, ru = min(r, d, 4 * H), min(r, H)
        if form == "singular":
            macs += rw * (d + 4 * H) + ru * 5 * H
        else:
            macs += rw * (d + 4 * H - rw) + ru * (5 * H - ru)
        d = H
    return 2 * macs


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": j["hbm_gbs"], "bf16_tflops": j["bf16_tflops"],
                "bf16_tflops_sustained": j.get("bf16_tflops_sustained", j["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe):
    one background `nvidia-smi -lms 50` process, started before and killed after the timed loops."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.t_start = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.3)    # let the first sample land before the timed region starts
        except Exception:
            self.proc = None

    def mark(self):
        """Call right before the timed region: samples taken earlier are dropped."""
        self.t_start = time.time()

    def stop(self):
        lines = []
        if self.proc is not None:
            time.sleep(0.06)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
                lines = [l for l in out.splitlines() if l.strip()]
            except Exception:
                self.proc.kill()
        sm, mx, reasons, pw = [], 0.0, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in lines:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
                pw.append(float(f[2]))
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        # the GPU is idle before mark(): keep the samples taken under load (upper half by power draw) for the median
        if len(sm) >= 4:
            order = np.argsort(pw)[len(pw) // 2:]
            sm_load = [sm[i] for i in order]
        else:
            sm_load = sm
        return {"sm_mhz": float(np.median(sm_load)) if sm_load else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


def build_workload(a, svdlstm):
    layers, dense = svdlstm.synthetic_layers(16, a.hidden, a.layers, seed=0)
    full = svdlstm.full_model_from_weights(layers, dense, return_sequences=True)
    sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=True, return_sequences=True)
    model = svdlstm.truncate_singular_model(sm, a.rank)
    model._singular_parent = sm
    model._full_parent = full
    return layers, dense, model


def cpu_oracle_model(layers, dense, rank):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import svdlstm_oracle as O
    ofull = O.model_from_weights(layers, dense, dtype=np.float32)
    osm = O.make_LSTM_singular_model(ofull, merged_kernel=True, return_sequences=True, svd_dtype=np.float32, dtype=np.float32)
    return O.truncate_singular_model(osm, rank, dtype=np.float32)


def cpu_forward_sample(om, Bs, Ts, seed=0):
    """Reference maths (oracle port, float32, batched numpy => multithreaded BLAS) on a bounded sample."""
    x = np.random.default_rng(seed).standard_normal((Bs, Ts, 16)).astype(np.float32)
    t0 = time.perf_counter()
    om.predict(x)
    dt = time.perf_counter() - t0
    return Bs * Ts / dt, dt


def cpu_sample_shape(om, a, target_s):
    """Sample = the full batch (same per-step matrix shapes as the GPU workload) for as many timesteps as fit
    `target_s` seconds of host time, found by timing 2 steps first."""
    Bs = a.batch if not a.quick else 64
    om.predict(np.zeros((Bs, 1, 16), np.float32))          # warm-up (BLAS threads, page faults)
    _, dt = cpu_forward_sample(om, Bs, 2)
    Ts = int(max(2, min(a.seq_len, target_s / max(dt / 2, 1e-6))))
    return Bs, Ts


def reference_arm(a):
    """--impl reference: the reference's maths (oracle port; TF/Keras itself is not installable here)
    on the host cores, bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import svdlstm_oracle as O
    layers, dense = O.synthetic_layers(16, a.hidden, a.layers, seed=0)
    om = cpu_oracle_model(layers, dense, a.rank)
    cores = os.cpu_count() or 1
    Bs, Ts = cpu_sample_shape(om, a, target_s=6.0 if not a.quick else 0.5)
    vals = []
    for i in range(a.warmup + a.steps):
        v, dt = cpu_forward_sample(om, Bs, Ts, seed=i)
        if i >= a.warmup:
            vals.append((v, dt))
    v = float(np.mean([x[0] for x in vals]))
    ms = float(np.mean([x[1] for x in vals])) * 1e3
    sample = "B=%d of %d sequences x T=%d of %d steps per step, float32 numpy oracle (multithreaded BLAS)" % (Bs, a.batch, Ts, a.seq_len)
    line = {"impl": "reference", "metric": "low-rank LSTM timesteps/sec (batch 4096)", "value": v, "unit": "sequence-timesteps/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a),
            "cpu_baseline": {"value": v, "unit": "sequence-timesteps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "sequence-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(a):
    return {"workload": "synthetic %d-layer SVD-LSTM (3-factor, merged) D=16 H=%d seq_len=%d batch=%d/GPU rank=%d + Dense(1)"
                        % (a.layers, a.hidden, a.seq_len, a.batch, a.rank),
            "rank": a.rank, "form": "singular (3-factor)", "batch_per_gpu": a.batch, "seq_len": a.seq_len, "hidden": a.hidden, "layers": a.layers,
            "l2": "inputs larger than L2 (x is %.0f MB per GPU)" % (a.batch * a.seq_len * 16 * 4 / 1e6),
            "parallelism": "independent sequences per GPU, no data-path collective"}


def batch1_table(svdlstm, torch):
    """configs[1]: shipped DROPBEAR model, batch 1, us/timestep at full rank and truncated ranks."""
    layers, dense = svdlstm.load_model_weights_npz(os.path.join(ROOT, "tests", "golden", "dropbear_weights.npz"))
    full = svdlstm.full_model_from_weights(layers, dense)
    sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=True, return_sequences=True)
    T = 20000
    x = torch.randn(1, T, 16, generator=torch.Generator().manual_seed(0)).cuda()
    out = {}

    def tm(model, label, engine=None):
        model(x[:, :256], engine=engine)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model(x, engine=engine)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[label] = round(best * 1e3 / T, 4)

    tm(full, "full")
    for r in (15, 12, 8, 4, 1):
        tm(svdlstm.truncate_singular_model(sm, r), "3F_r%d" % r)
        tm(svdlstm.make_LSTM_reduced_model(sm, rank=r), "2F_r%d" % r)
    tm(full, "full_general_engine", engine="general")
    # True streaming (the reference's real-time setting, svd_acceleration_v3.py:151: one sample every 400 us): one call per chunk of
    # `c` samples through Sequential.stream with the state carried on the device -- pinned host sample in, host prediction out.
    import time
    import numpy as np
    streaming = {}
    m8 = svdlstm.truncate_singular_model(sm, 8)
    for label, model in (("full", full), ("3F_r8", m8)):
        for c in (1, 16):
            xin = torch.randn(1, c, 16).pin_memory()
            yout = torch.empty(1, c, 1).pin_memory()
            state = None
            lat = []
            for it in range(300):
                t0 = time.perf_counter()
                xd = xin.cuda(non_blocking=True)
                y, state = model.stream(xd, state)
                yout.copy_(y, non_blocking=True)
                torch.cuda.synchronize()
                lat.append(time.perf_counter() - t0)
            lat = np.array(lat[50:]) * 1e6
            streaming["%s_chunk%d" % (label, c)] = {"call_us_median": round(float(np.median(lat)), 1), "call_us_p99": round(float(np.percentile(lat, 99)), 1),
                                                    "us_per_sample": round(float(np.median(lat)) / c, 1)}
    out_streaming = {"how": "Sequential.stream per chunk: pinned H2D of the chunk, one launch with state in/out, D2H of the predictions, "
                            "synchronize; wall clock of the Python call, 250 calls after 50 warm-ups", "calls": streaming}
    # algorithmic on-chip bytes/step of the 3-factor model at full rank: 28 140 B
    byt = 28140.0
    return {"unit": "us/timestep", "T": T, "model": "DROPBEAR 3x15 LSTM + Dense(1), batch 1", "us_per_step": out,
            "onchip_roofline": {"bytes_per_step_3F_r15": byt, "achieved_gbs": round(byt / (out["3F_r15"] * 1e-6) / 1e9, 2),
                                "peak_gbs_1sm": 251.5, "frac": round(byt / (out["3F_r15"] * 1e-6) / 1e9 / 251.5, 4),
                                "note": "peak = 128 B/clk x 1965 MHz, one SM (the latency chain uses one CTA)"},
            "streaming": out_streaming, "realtime_budget_us": 400.0}


def main():
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
        return
    import torch
    import svdlstm
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (native) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    layers, dense, model = build_workload(a, svdlstm)
    smodel = model._singular_parent
    B, T, D = a.batch, a.seq_len, 16
    gen = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(B, T, D, generator=gen).pin_memory()
    x = x_host.to(dev, non_blocking=True)
    torch.cuda.synchronize()
    engine = a.engine

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    torch.cuda.profiler.start()        # ncu --profile-from-start off: profile the warm-up + timed forwards only (not the model build)
    for _ in range(max(a.warmup, 3)):
        y = model(x, engine=engine)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = svdlstm.launches()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for i in range(a.steps):
        evs[i][0].record()
        y = model(x, engine=engine)
        evs[i][1].record()
    t1.record()
    barrier()
    launches = svdlstm.launches() - l0
    torch.cuda.profiler.stop()
    total_ms = t0.elapsed_time(t1)
    kern_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in evs]))
    tt = torch.tensor([total_ms, kern_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, kern_ms = float(tt[0]), float(tt[1])
    ms_per_step = total_ms / a.steps
    value = world * B * T / (ms_per_step * 1e-3)

    # ---- end to end through the public API: pinned host X -> device -> forward -> y back to host -----
    # A serving loop: every step copies ITS OWN input from pinned host memory and reads its result back, all inside the
    # timed region; the copy of step i+1 (copy stream, second device buffer) overlaps the forward of step i.
    e2e = None
    if not a.no_e2e:
        y_host = [torch.empty((B, T, 1), dtype=torch.float32).pin_memory() for _ in range(2)]
        x_dev = [torch.empty((B, T, D), dtype=torch.float32, device=dev) for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        main_stream = torch.cuda.current_stream(dev)
        h2d_done = [torch.cuda.Event() for _ in range(2)]
        x_free = [torch.cuda.Event() for _ in range(2)]

        def serve(n_steps):
            for b2 in range(2):
                x_free[b2].record(main_stream)
            with torch.cuda.stream(copy_stream):            # prologue: input of step 0
                x_dev[0].copy_(x_host, non_blocking=True)
                h2d_done[0].record(copy_stream)
            for i in range(n_steps):
                cur, nxt = i & 1, (i + 1) & 1
                if i + 1 < n_steps:
                    with torch.cuda.stream(copy_stream):    # input of step i+1 while step i computes
                        copy_stream.wait_event(x_free[nxt])
                        x_dev[nxt].copy_(x_host, non_blocking=True)
                        h2d_done[nxt].record(copy_stream)
                main_stream.wait_event(h2d_done[cur])
                yy = model(x_dev[cur], engine=engine)
                x_free[cur].record(main_stream)
                y_host[cur].copy_(yy, non_blocking=True)    # result of step i back to pinned host memory

        serve(2)
        barrier()
        s0 = torch.cuda.Event(enable_timing=True)
        s1 = torch.cuda.Event(enable_timing=True)
        s0.record()
        serve(a.steps)
        s1.record()
        barrier()
        te = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_ms = float(te[0]) / a.steps
        e2e = {"value": world * B * T / (e2e_ms * 1e-3), "unit": "sequence-timesteps/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(x_host.numel() * 4), "d2h_bytes_per_step": int(y_host[0].numel() * 4),
               "how": "Sequential.__call__ per step; H2D of step i+1 on a copy stream overlaps the forward of step i; y copied back every step"}

    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    pk = peaks()
    fl = flops_per_seq_step(D, a.hidden, a.layers, a.rank) * B * T
    achieved = fl / (kern_ms * 1e-3) / 1e12
    eng_id = model.last_engine()
    eng_name = {0: "auto", 1: "general(fp32 cuda cores)", 2: "wavefront(fp32)", 3: "tc(tcgen05 f16 operands, fp32 accumulate)"}.get(eng_id, str(eng_id))
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")     # dram bytes of one forward, from the committed ncu --set full capture
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("rank_%d" % a.rank)
        except Exception:
            traffic = None
    # rows handed from layer to layer per sequence-step: the next layer's t_w (rank rows, padded to 16) when they fit the S1u
    # TMEM tiles next to this layer's t_u rows, else the hidden state itself
    r_eff = min(a.rank, a.hidden)
    handoff_rows = (r_eff + 15) // 16 * 16 if 2 * ((r_eff + 7) // 8 * 8) <= 256 else a.hidden
    roofline = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_tflops_sustained"], "traffic": traffic,
                "peak_source": pk["source"] + " (cuBLAS bf16 sustained; f16 runs at the same tensor rate)",
                "kernel": "lstm_tc_pipe_kernel: all %d layers in one co-resident launch, 64-sequence tiles (one forward = pack_x + this launch)"
                          % a.layers if eng_id == 3 else eng_name,
                "algorithmic_flops_per_launch": fl, "kernel_ms": kern_ms,
                "hbm_bytes_algorithmic": int(B * T * (D * 4 + 4) + (a.layers - 1) * 2 * B * T * handoff_rows * 2 + B * T * D * 2 * 2)}
    line = {"metric": "low-rank LSTM timesteps/sec (batch 4096)", "value": value, "unit": "sequence-timesteps/s", "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16" if eng_id == 3 else "f32", "data": "synthetic", "config": workload_config(a),
            "engine": eng_name, "roofline": roofline, "clocks": clocks, "gpu_launches": launches}
    if eng_id == 3:
        # reduced-precision report (north_star): RMSE of the tensor-core output against the FP32 parity engine, same weights/inputs
        xs = x[:256].contiguous()
        y32 = model(xs, engine="general")
        ytc = model(xs, engine=engine)
        line["rmse_delta"] = {"rmse_tc_vs_fp32": float(((ytc - y32) ** 2).mean().sqrt()), "max_abs": float((ytc - y32).abs().max()),
                              "output_rms": float((y32 ** 2).mean().sqrt()),
                              "rmse_last_64_steps": float(((ytc[:, -64:] - y32[:, -64:]) ** 2).mean().sqrt()),
                              "sample": "first 256 sequences x all %d steps" % T}
    if not a.no_sweep and world == 1 and eng_id == 3:
        sweep = {}
        for r in (8, 16, 32, 64, 128, 256):
            m = svdlstm.truncate_singular_model(smodel, r)
            for _ in range(2):
                m(x, engine=engine)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                m(x, engine=engine)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 2
            tf = flops_per_seq_step(D, a.hidden, a.layers, r) * B * T / (ms * 1e-3) / 1e12
            sweep["r%d" % r] = {"ms": round(ms, 3), "Mseqsteps_per_s": round(B * T / ms / 1e3, 1), "tflops": round(tf, 1),
                                "frac_of_peak": round(tf / pk["bf16_tflops_sustained"], 4)}
        if a.hidden <= 256:
            # the uncompressed LSTM on the same engine (runs as the factorisation I . W): what the reference's speed-up plots divide by
            fm = model._full_parent
            for _ in range(2):
                fm(x, engine=engine)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                fm(x, engine=engine)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 2
            d_in, full_macs = D, 0
            for _ in range(a.layers):
                full_macs += d_in * 4 * a.hidden + a.hidden * 4 * a.hidden
                d_in = a.hidden
            tf = 2 * full_macs * B * T / (ms * 1e-3) / 1e12
            sweep["full"] = {"ms": round(ms, 3), "Mseqsteps_per_s": round(B * T / ms / 1e3, 1), "tflops": round(tf, 1),
                             "frac_of_peak": round(tf / pk["bf16_tflops_sustained"], 4),
                             "note": "uncompressed LSTM (4(DH+H^2) MACs per layer-step) on the same tensor-core engine"}
            for r in (8, 16, 32, 64, 128, 256):
                sweep["r%d" % r]["speedup_vs_full"] = round(ms / sweep["r%d" % r]["ms"], 2)
        line["rank_sweep"] = sweep
    if e2e is not None:
        line["e2e"] = e2e
    if not a.no_batch1 and world == 1:
        line["batch1_us_per_step"] = batch1_table(svdlstm, torch)
    if not a.no_cpu_baseline and world == 1:
        om = cpu_oracle_model(layers, dense, a.rank)
        Bs, Ts = cpu_sample_shape(om, a, target_s=12.0 if not a.quick else 0.5)
        v, dt = cpu_forward_sample(om, Bs, Ts)
        line["cpu_baseline"] = {"value": v, "unit": "sequence-timesteps/s", "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": "B=%d of %d sequences x T=%d of %d steps, float32 numpy oracle (multithreaded BLAS), %.1f s"
                                          % (Bs, B, Ts, T, dt)}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
