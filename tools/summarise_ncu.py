"""Summarises ncu outputs into the small text files kept under profiles/.
    python tools/summarise_ncu.py launches gpurun_out/launches_X.csv  > profiles/launches_X.md
    python tools/summarise_ncu.py full gpurun_out/prof_X.ncu-rep       > profiles/prof_X.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    g = hdr.index("Grid Size")
    agg = OrderedDict()
    for r in rows:
        name = r[k].replace("void ", "").replace("vsr::", "").split("(")[0]
        a = agg.setdefault(name, [0, 0.0, r[g]])
        a[0] += 1
        a[1] += float(r[v].replace(",", "")) / 1e3
    tot = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | total us | share | avg us | grid |\n|---|---|---|---|---|---|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {a[0]} | {a[1]:.0f} | {a[1] / tot:.3f} | {a[1] / a[0]:.1f} | {a[2]} |")
    print(f"\n{len(rows)} launches, {tot / 1e3:.2f} ms summed (cold-cache, serialised: compare shares, not absolutes)")


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "smsp__cycles_active.avg"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, rows = rows[0], rows[1], rows[2:]
    cols = {}
    for key in KEYS + ["sm__pipe_tensor_cycles_active", "tensor"]:
        for i, h in enumerate(hdr):
            if key in h and h not in cols:
                if key in KEYS or len([c for c in cols if key in c]) < 6:
                    cols[h] = i
    k = hdr.index("Kernel Name")
    for r in rows:
        print(f"### {r[hdr.index('ID')]}: `{r[k].replace('void ', '').replace('vsr::', '').split('(')[0]}` grid {r[hdr.index('Grid Size')]}")
        for h, i in cols.items():
            print(f"- {h} [{units[i]}]: {r[i]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
