#!/usr/bin/env python3
"""Per-source-line instruction and stall-sample shares of one kernel of an ncu report captured with
--import-source on.   python tools/ncu_lines.py report.ncu-rep <kernel id> [top N]"""
import csv
import subprocess
import sys

rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", ":::" + kid],
                     capture_output=True, text=True).stdout
cur, hdr, agg = None, None, {}
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]
        continue
    if r[0] == "Function Name":
        print(r[1][:100])
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 8 and r[2] == "-":
        try:
            ie, smp = int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        if ie > 0 or smp > 0:
            agg[(cur.split("/")[-1], int(r[0]))] = (ie, smp, r[1].strip()[:100])
tot, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print("instructions executed %d, samples %d" % (tot, ts))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-22s %4d %6.2f%% inst %6.2f%% smp  %s" % (k[0], k[1], 100.0 * v[0] / tot, 100.0 * v[1] / max(ts, 1), v[2]))
