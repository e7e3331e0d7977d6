"""Instruction mix and hottest SASS lines of one kernel of an ncu report (needs --import-source on).
    python tools/ncu_source.py report.ncu-rep <launch index> [top]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][:2])
hdr = rows[1]
rows = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index("Instructions Executed")].isdigit() and r[0].startswith("0x")]
ie, ss, src = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
tot = sum(int(r[ie]) for r in rows)
tots = max(1, sum(int(r[ss]) for r in rows))
print("total warp-instructions", tot, "samples", tots, "sass lines", len(rows))
c, cs = Counter(), Counter()
for r in rows:
    toks = r[src].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    c[op] += int(r[ie])
    cs[op] += int(r[ss])
for op, n in c.most_common(22):
    print(f"{op:16s} {n:10d} {n / tot:6.3f}  stall-samples {cs[op] / tots:6.3f}")
print()
for r in sorted(rows, key=lambda r: -int(r[ss]))[:top]:
    print(f"{int(r[ss]):7d} {int(r[ie]):9d}  {r[src][:120]}")
