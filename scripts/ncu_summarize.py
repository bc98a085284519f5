"""Summarise gpurun_out/launches.csv (ncu launch list) and gpurun_out/full_*.ncu-rep (ncu --set full) into the text
files committed under profiles/.   usage: python scripts/ncu_summarize.py <tag>      (e.g. r01_final)"""
import collections
import csv
import glob
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
CMD = "python bench.py --no-graph --steps 1 --warmup 0 --min-warmup 0 --no-e2e --no-cpu-baseline"

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size"]


def launch_summary():
    path = os.path.join(ROOT, "gpurun_out", "launches.csv")
    rows = [l for l in open(path, errors="replace") if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v_ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        t = tot[r["Kernel Name"]]
        t[0] += 1
        t[1] += v_ms
    total = sum(v[1] for v in tot.values())
    n = sum(v[0] for v in tot.values())
    out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 ; {CMD}",
           f"# eager training steps (roofline-timing step(s) + 1 timed), B=64, Swin-T W7 2+8 crops DDINO K=65536; "
           f"{n} launches, {total:.1f} ms summed (cold-cache, serialised)",
           "# share%  launches  total_ms  kernel"]
    for k, (c, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
        out.append(f"{100 * ms / total:6.2f} {c:6d} {ms:9.3f}  {k[:110]}")
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_launch_summary.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:16]))


def full_metrics():
    out = [f"# ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <n> -c 1 ; {CMD}"]
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "full_*.ncu-rep"))):
        name = os.path.basename(rep)[5:-8]
        r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(r.stdout)))
        if len(rows) < 3:
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        out.append(f"== {name}")
        out.append(f"   Kernel Name: {d.get('Kernel Name', ('?',))[0][:160]}")
        for k in KEYS:
            if k in d:
                out.append(f"   {k}: {d[k][0]} {d[k][1]}")
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_key_metrics.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def kernel_cases():
    """gpurun_out/k_<case>.ncu-rep (scripts/ncu_capture_r2.sh) -> profiles/<tag>_ncu_kernels.txt + profiles/ncu_traffic.json
    (what bench.py's roofline.traffic reads: measured DRAM bytes of the launch next to its algorithmic bytes)."""
    import json
    cases = json.load(open(os.path.join(ROOT, "gpurun_out", "ncu_cases.json")))
    try:
        commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    except Exception:
        commit = "?"
    out = ["# ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <n> -c 1 python scripts/ncu_kernels.py <case>",
           "# one launch of each hot kernel at a known shape of the Swin-T B=64 step; algorithmic bytes / flops from scripts/ncu_kernels.py"]
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic = json.load(open(tpath)) if os.path.isfile(tpath) else {}   # captures of other cases are kept
    only = [a.split("=", 1)[1].split(",") for a in sys.argv if a.startswith("--cases=")]
    for case, meta in cases.items():
        if only and case not in only[0]:
            continue
        rep = os.path.join(ROOT, "gpurun_out", f"k_{case}.ncu-rep")
        csvp = os.path.join(ROOT, "gpurun_out", f"k_{case}.csv")
        if os.path.isfile(csvp) and os.path.getsize(csvp) > 100:
            text = open(csvp, errors="replace").read()
        elif os.path.isfile(rep):
            text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        else:
            continue
        rows = list(csv.reader(io.StringIO(text)))
        if len(rows) < 3:
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}

        def num(key):
            if key not in d:
                return None
            v, u = d[key]
            x = float(v.replace(",", ""))
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3,
                    "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}.get(u, 1.0)
            return x * mult
        dram = (num("dram__bytes_read.sum") or 0.0) + (num("dram__bytes_write.sum") or 0.0)
        dur = num("gpu__time_duration.sum")
        out.append(f"== {case}: {meta['note']}")
        out.append(f"   Kernel Name: {d.get('Kernel Name', ('?',))[0][:160]}")
        for k in KEYS:
            if k in d:
                out.append(f"   {k}: {d[k][0]} {d[k][1]}")
        out.append(f"   algorithmic bytes {meta['algorithmic_bytes'] / 1e6:.1f} MB, measured DRAM traffic {dram / 1e6:.1f} MB "
                   f"(x{dram / max(meta['algorithmic_bytes'], 1):.2f}); under ncu: {dur * 1e6:.1f} us = "
                   f"{meta['algorithmic_bytes'] / dur / 1e9:.0f} GB/s algorithmic"
                   + (f", {meta['algorithmic_flops'] / dur / 1e12:.0f} TFLOP/s" if meta["algorithmic_flops"] else ""))
        traffic[case] = {"ncu_launch": f"{d.get('Kernel Name', ('?',))[0][:80]} | {meta['note']}", "dram_bytes": dram,
                         "algorithmic_bytes": meta["algorithmic_bytes"], "commit": commit}
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_kernels.txt"), "w").write("\n".join(out) + "\n")
    # keys bench.py looks up
    alias = {"window_attn_bwd": "attn_bwd7_s0", "window_attn_fwd": "attn_fwd7_s0", "dino_ce_bwd": "dino_ce_bwd",
             "gemm_bias_act": "gemm_gelu_fc1_0", "gemm_mul_colsum": "gemm_mul_fc2dgrad_0", "gemm_bf16": "gemm_fwd_qkv0",
             "gemm_wgrad": "gemm_wgrad_qkv0", "add_ln_bwd": "add_ln_bwd_96", "add_ln_fwd": "add_ln_fwd_96"}
    for k, c in alias.items():
        if c in traffic:
            traffic[k] = traffic[c]
    json.dump(traffic, open(tpath, "w"), indent=1)
    print("\n".join(out))


if __name__ == "__main__":
    if os.path.isfile(os.path.join(ROOT, "gpurun_out", "launches.csv")) and "--no-list" not in sys.argv:
        launch_summary()
    if "--kernels" in sys.argv:
        kernel_cases()
    else:
        full_metrics()
