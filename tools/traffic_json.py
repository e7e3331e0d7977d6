"""ncu csv (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum) -> the per-launch json
bench.py reads for roofline.traffic.
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \\
        --log-file gpurun_out/traffic_X.csv python tools/profile_srfbn.py 20 270 480 1
    python tools/traffic_json.py gpurun_out/traffic_X.csv > profiles/traffic_X.json"""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    iid, k, m, u, v = (hdr.index(c) for c in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    out = {}
    for r in rows:
        e = out.setdefault(int(r[iid]), {"id": int(r[iid]), "kernel": r[k].replace("void ", "").replace("vsr::", "").split("(")[0]})
        val = float(r[v].replace(",", "")) * UNIT.get(r[u], 1.0)
        if r[m] == "dram__bytes_read.sum":
            e["dram_read_bytes"] = val
        elif r[m] == "dram__bytes_write.sum":
            e["dram_write_bytes"] = val
        elif r[m] == "gpu__time_duration.sum":
            e["us"] = val
    print(json.dumps([out[i] for i in sorted(out)], indent=0))


if __name__ == "__main__":
    main(sys.argv[1])
