#!/usr/bin/env python3
"""Per-unit figures of the motion-search kernels for bench.py's roofline lines, from an `ncu --set full
--import-source on` report: executed thread-instructions per luma pixel (of the pyramid level) per reference, split by
what issues them, and the DRAM bytes per reference.

    python tools/ncu_units.py report.ncu-rep [out.json]

Classes (SASS opcode prefix):  int = the integer arithmetic the roofline is about (fma and alu pipes: IMAD, IADD3,
VIADD, LOP3, SHF, IDP, PRMT, IABS, ISETP, SEL, VIMNMX, LEA, I2IP, ...), mem = loads/stores/atomics, shfl = shuffles and
votes, ctl = branches, barriers, convergence and scheduling, uni = uniform datapath, other = the rest.
"""
import collections
import csv
import json
import subprocess
import sys

MODEL = {
    # minimal integer operations of the kernels' own de-duplicated formulations, per pixel of the searched level and
    # per reference (derivation: DESIGN.md section 4)
    "luma_search_2step": 394.0,
    "luma_search_1step": 210.0,
}


def classify(op):
    base = op.split(".")[0]
    if base.startswith("U") and base not in ("UNPACK",):
        return "uni"
    if base in ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDL", "STL", "ATOMS", "ATOMG", "ATOM", "RED", "LDC", "LDSM", "LDGSTS", "CCTL", "MEMBAR", "ERRBAR", "FENCE"):
        return "mem"
    if base in ("SHFL", "VOTE", "VOTEU", "MATCH", "REDUX"):
        return "shfl"
    if base in ("BRA", "BRX", "JMP", "EXIT", "RET", "CALL", "BAR", "BSSY", "BSYNC", "WARPSYNC", "NOP", "DEPBAR", "YIELD", "NANOSLEEP", "BREAK", "BMOV",
                "ELECT", "SYNCS", "UTMALDG", "ACQBULK", "ENDCOLLECTIVE", "KILL", "BPT"):
        return "ctl"
    if base in ("IMAD", "IADD3", "IADD", "VIADD", "LOP3", "LOP", "SHF", "SHL", "SHR", "IDP", "IDP4A", "PRMT", "IABS", "ISETP", "SEL", "VIMNMX", "VIMNMX3", "IMNMX",
                "LEA", "I2IP", "VABSDIFF", "VABSDIFF4", "BMSK", "SGXT", "FLO", "POPC", "MOV", "PLOP3", "CS2R", "S2R", "R2UR", "S2UR", "P2R", "R2P", "ISCADD", "IMUL",
                "I2I", "IADD32I", "LOP32I"):
        return "int"
    return "other"


def kernels(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, un = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    res = []
    for r in rows[2:]:
        d = {"id": r[h.index("ID")], "name": r[h.index("Kernel Name")].split("(")[0].replace("vp8::", ""),
             "grid": [int(x) for x in r[h.index("Grid Size")].strip("() ").split(",")],
             "us": float(r[h.index("gpu__time_duration.sum")]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(un[h.index("gpu__time_duration.sum")], 1e-3),
             "dram": sum(float(r[h.index(m)]) * scale.get(un[h.index(m)], 1.0) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))}
        for m, key in (("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
                       ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_pct"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma_pct"),
                       ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier_per_issue"), ("launch__registers_per_thread", "registers")):
            if m in h:
                d[key] = float(r[h.index(m)])
        res.append(d)
    return res


def instruction_mixes(rep):
    """one Counter per result of the report, in report order (the source page lists them one after the other)"""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    hdr, mixes, names = None, [], []
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "Kernel Name":
            mixes.append(collections.Counter())
            names.append(r[1])
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and mixes and len(r) == len(hdr):
            try:
                n = int(r[hdr.index("Thread Instructions Executed")])
            except ValueError:
                continue
            toks = r[hdr.index("Source")].split()
            op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
            mixes[-1][classify(op)] += n
    # (a result whose code spans two source files is listed once per file, with the same numbers: keep one)
    keep = [i for i in range(len(mixes)) if i == 0 or not (names[i] == names[i - 1] and mixes[i] == mixes[i - 1])]
    return [mixes[i] for i in keep]


def main(rep, out_json=None):
    res = {}
    ks = kernels(rep)
    mixes = instruction_mixes(rep)
    assert len(mixes) == len(ks), (len(mixes), len(ks))
    for k, mix in zip(ks, mixes):
        if k["name"] not in ("k_luma_search_2step", "k_luma_search_1step"):
            continue
        name = k["name"][2:]
        refs = k["grid"][1]
        pixels = (1920 * 1088 if name == "luma_search_2step" else k["grid"][0] * 8 * 64) * refs
        key = name if name == "luma_search_2step" else "%s@%d" % (name, k["grid"][0])
        total = float(sum(mix.values()))
        e = res.setdefault(key, {"captures": 0, "thread_instr": 0.0, "int": 0.0, "pixels_x_refs": 0.0, "dram": 0.0, "refs": 0, "us": 0.0,
                                 "mix": collections.Counter(), "sample": {}})
        e["captures"] += 1
        e["thread_instr"] += total
        e["int"] += mix["int"]
        e["mix"].update(mix)
        e["pixels_x_refs"] += pixels
        e["dram"] += k["dram"]
        e["refs"] += refs
        e["us"] += k["us"]
        e["sample"] = {x: k[x] for x in ("issue_active_pct", "warps_active_pct", "pipe_alu_pct", "pipe_fma_pct", "stall_barrier_per_issue", "registers", "grid") if x in k}
    final = {"source": rep.split("/")[-1]}
    # all full-resolution launches of luma_search_1step stand for the kernel (the per-pixel figure is the same at every level)
    for key, e in res.items():
        name = key.split("@")[0]
        d = {"captures": e["captures"], "thread_instr_per_pixel_per_ref": e["thread_instr"] / e["pixels_x_refs"],
             "alu_pipe_thread_instr_per_pixel_per_ref": e["int"] / e["pixels_x_refs"],
             "instruction_mix": {c: round(n / e["thread_instr"], 4) for c, n in sorted(e["mix"].items())},
             "min_ops_per_pixel_per_ref": MODEL[name], "dram_bytes_per_ref": e["dram"] / e["refs"],
             "ncu_us_per_launch": e["us"] / e["captures"], "ncu": e["sample"]}
        final[key] = d
    big = [k for k in final if k.startswith("luma_search_1step@")]
    if big:
        best = max(big, key=lambda k: int(k.split("@")[1]))
        final["luma_search_1step"] = final[best]
    print(json.dumps(final, indent=1))
    if out_json:
        json.dump(final, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:])
