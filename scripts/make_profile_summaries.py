#!/usr/bin/env python
"""Turns the raw captures of scripts/profile_round.sh (gpurun_out/, scratch) into the tracked evidence under profiles/:

    python scripts/make_profile_summaries.py r02 [commit-of-the-capture]

  - profiles/launches_bench_<R>.csv / launches_c5_<R>.csv     the ncu launch lists, copied
  - profiles/launch_shares_<R>.txt                            per-kernel share of one forward + DRAM bytes
  - profiles/ncu_traffic.json                                 DRAM bytes per forward + tensor-pipe figures, read by bench.py's roofline
  - profiles/tc_pipe_rank{256,128,32}_<R>.txt, tc_pipe_c5_<R>.txt, b1_wavefront_<R>.txt     ncu --set full condensed (ncu_summary.py)
  - profiles/sass_histogram_<R>.txt                           cuobjdump -sass opcode histogram of libsvdlstm.so (UTCHMMA / LDTM / UBLKCP ...)
"""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
O = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def head():
    try:
        return subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    except Exception:
        return "?"


def launch_table(path):
    rows = [l for l in open(path) if l.startswith('"')]
    rd = list(csv.DictReader(io.StringIO("".join(rows))))
    per = collections.OrderedDict()
    for r in rd:
        k = (r["ID"], r["Kernel Name"])
        per.setdefault(k, {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9,
                                                                                               "ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(r["Metric Unit"], 1)
    return per


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").replace("svdlstm::", "").replace("unnamed>::", "").replace("<unnamed>::", "")[:60]


def shares(per, n_forwards, out):
    agg = collections.OrderedDict()
    for (_, name), m in per.items():
        a = agg.setdefault(short(name), [0, 0.0, 0.0])
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0)
        a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    for k, a in agg.items():
        out.write("%-48s n=%3d  %8.3f ms/forward  %5.1f%% of the step  dram %8.1f MB/forward\n" % (k, a[0], a[1] / 1e6 / n_forwards, 100 * a[1] / tot, a[2] / 1e6 / n_forwards))
    return agg, tot


def raw_metric(rep, metric, pat):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return None
    hdr = rows[0]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        if pat in r[ki] and metric in hdr:
            try:
                return float(r[hdr.index(metric)].replace(",", ""))
            except ValueError:
                return None
    return None


def main():
    os.makedirs(P, exist_ok=True)
    # the commit the captures were taken at (gpurun snapshots have no .git): second argument, else the current HEAD
    stamp = "commit %s" % (sys.argv[2] if len(sys.argv) > 2 else head())
    traffic = {"unit": "bytes per forward", "stamp": stamp}
    lb = os.path.join(O, "launches_%s.csv" % R)
    if os.path.exists(lb):
        shutil.copy(lb, os.path.join(P, "launches_bench_%s.csv" % R))
        per = launch_table(lb)
        n_fw = sum(1 for (_, n) in per if "lstm_tc_pipe_kernel" in n) or sum(1 for (_, n) in per if "pack_x_kernel" in n) or 1
        with open(os.path.join(P, "launch_shares_%s.txt" % R), "w") as f:
            f.write("# %s; bench.py --steps 3 --warmup 3 (profiled region = warm-up + timed forwards = %d forwards), C3 rank 128; ncu launch list: cold-cache, serialised\n" % (stamp, n_fw))
            agg, tot = shares(per, n_fw, f)
        fw = sum(a[2] for k, a in agg.items() if "pack_x" in k or "lstm_tc" in k) / n_fw
        traffic["rank_128"] = fw
        traffic["source"] = ("profiles/launches_bench_%s.csv (%s): dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of one forward "
                             "(lstm_tc_pipe_kernel, plus pack_x when the build uses it), C3 rank 128, mean of %d forwards" % (R, stamp, n_fw))
        tc = sum(a[1] for k, a in agg.items() if "lstm_tc" in k)
        traffic["kernel_share_of_step_rank_128"] = tc / tot
    lc = os.path.join(O, "launches_c5_%s.csv" % R)
    if os.path.exists(lc):
        shutil.copy(lc, os.path.join(P, "launches_c5_%s.csv" % R))
    summ = os.path.join(ROOT, "scripts", "ncu_summary.py")
    for rep, pat, dst in [("tc_layer_rank256_%s" % R, "lstm_tc", "tc_pipe_rank256_%s.txt" % R), ("tc_layer_rank128_%s" % R, "lstm_tc", "tc_pipe_rank128_%s.txt" % R),
                          ("tc_layer_rank32_%s" % R, "lstm_tc", "tc_pipe_rank32_%s.txt" % R), ("tc_pipe_c5_%s" % R, "lstm_tc", "tc_pipe_c5_%s.txt" % R),
                          ("b1_wavefront_%s" % R, "lstm_wavefront", "b1_wavefront_%s.txt" % R),
                          ("general_fp32_%s" % R, "lstm_general", "general_fp32_%s.txt" % R)]:
        rp = os.path.join(O, rep + ".ncu-rep")
        pre = os.path.join(O, rep + ".txt")          # condensed on the GPU box by profile_round.sh (the reports exceed gpurun's 64 MiB)
        if os.path.exists(pre) and os.path.getsize(pre) > 200:
            txt = open(pre).read()
        elif os.path.exists(rp):
            txt = subprocess.run([sys.executable, summ, rp, pat], capture_output=True, text=True).stdout
        else:
            continue
        hot = os.path.join(O, rep + ".hot.txt")
        if os.path.exists(hot) and os.path.getsize(hot) > 50:
            txt += "\n## hottest source lines (ncu --page source: warp stall samples per line, needs -lineinfo)\n" + open(hot).read()
        open(os.path.join(P, dst), "w").write("# %s\n" % stamp + txt)
        m = re.search(r"rank(\d+)", rep)
        if m:
            for key, metric in (("tensor_pipe_active_pct_rank_%s", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                                ("tensor_pipe_elapsed_pct_rank_%s", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")):
                mm = re.search(re.escape(metric) + r"\s+([0-9.]+)", txt)
                if mm:
                    traffic[key % m.group(1)] = float(mm.group(1))
    b1 = os.path.join(P, "b1_wavefront_%s.txt" % R)
    if os.path.exists(b1):
        txt = open(b1).read()
        wf = re.search(r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum\s+([0-9.]+)", txt)
        du = re.search(r"gpu__time_duration\.sum\s+([0-9.]+)\s+(\w+)", txt)
        if wf and du:
            t_s = float(du.group(1)) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(du.group(2), 1e-6)
            steps = 4096      # scripts/prof_batch1.py default T
            traffic["b1_wavefront"] = {"shared_wavefronts_per_step": float(wf.group(1)) / steps, "shared_gbs": float(wf.group(1)) * 128 / t_s / 1e9,
                                       "source": "profiles/b1_wavefront_%s.txt (%s): l1tex__data_pipe_lsu_wavefronts_mem_shared.sum x 128 B / gpu__time_duration, T=4096" % (R, stamp)}
    json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
    # SASS opcode histogram of the shipped library
    lib = os.path.join(ROOT, "lstm-acceleration-with-singular-value-decomposition_b200", "libsvdlstm.so")
    if os.path.exists(lib):
        sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
        hist = collections.Counter()
        per_fn = collections.defaultdict(collections.Counter)
        fn = "?"
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                fn = m.group(1)
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
            if m:
                op = m.group(1).split(".")[0]
                hist[op] += 1
                per_fn[fn][op] += 1
        with open(os.path.join(P, "sass_histogram_%s.txt" % R), "w") as f:
            f.write("# library built at commit %s (+ working tree); cuobjdump -sass libsvdlstm.so (sm_100a), opcode counts over all kernels\n" % head())
            key = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "MUFU", "FFMA2", "FFMA", "HMMA", "LDGSTS", "DFMA", "REDUX", "SHFL"]
            f.write("blackwell-native markers: " + ", ".join("%s=%d" % (k, hist.get(k, 0)) for k in key) + "\n\n")
            for op, n in hist.most_common(60):
                f.write("%-12s %8d\n" % (op, n))
            f.write("\n# per kernel (markers only)\n")
            for fn_, c in sorted(per_fn.items()):
                marks = {k: c[k] for k in key if c.get(k)}
                if marks:
                    dem = subprocess.run(["cu++filt", fn_], capture_output=True, text=True).stdout.strip() or fn_
                    f.write("%s\n    %s\n" % (short(dem), marks))


if __name__ == "__main__":
    main()
