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


if __name__ == "__main__":
    launch_summary()
    full_metrics()
