"""Dev: summaries for profiles/ from ncu output.
  python scripts/summarize_ncu.py launches <launches.csv> <steps>   -> per-kernel share of the LAST step
  python scripts/summarize_ncu.py full <report.ncu-rep>             -> key metrics per profiled kernel (needs ncu)"""
import csv, io, subprocess, sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__block_size", "launch__cluster_dim_x", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]


def launches(path, steps):
    rows = [r for r in csv.DictReader(l for l in open(path) if l.startswith('"'))]
    names = [r["Kernel Name"] for r in rows]
    first = next(i for i, n in enumerate(names) if "prep_x_kernel" in n)
    per = (len(rows) - first) // steps if steps > 1 else len(rows) - first
    # the last step starts at the last prep_x_kernel
    last = max(i for i, n in enumerate(names) if "prep_x_kernel" in n)
    sel = rows[last:]
    agg = OrderedDict()
    for r in sel:
        k = r["Kernel Name"][:100]
        us = float(r["Metric Value"]) / 1e3
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print("launches         us  share  kernel")
    for k, (n, us) in agg.items():
        print(f"{n:8d} {us:10.1f} {100 * us / tot:5.1f}%  {k}")
    print(f"{sum(a[0] for a in agg.values()):8d} {tot:10.1f} 100.0%  total")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    for row in rd[2:]:
        d = dict(zip(hdr, row))
        print("----", d.get("Kernel Name", "?")[:110])
        for k in KEYS:
            if k in d:
                print(f"  {k} = {d[k]} {units[hdr.index(k)]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 1)
    else:
        full(sys.argv[2])
