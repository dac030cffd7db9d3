#!/usr/bin/env python
"""Per-source-line view of an .ncu-rep (needs -lineinfo + --import-source on): warp instructions executed,
stall samples and the dominant stall reasons for every CUDA source line that executed anything."""
import csv, io, subprocess, sys

rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname, hdr, kern, first = None, None, None, None
lines = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        kern = r[1]
        if first is None:
            first = kern
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and kern == first:
        try:
            ie = int(r[hdr.index("Instructions Executed")])
            sm = int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        stalls = {}
        for i, n in enumerate(hdr):
            if n.startswith("stall_") and "Not Issued" not in n:
                try:
                    v = int(r[i])
                except ValueError:
                    v = 0
                if v:
                    stalls[n[6:]] = v
        lines.append((fname, int(r[0]), r[1].strip(), ie, sm, stalls))
ti = sum(l[3] for l in lines) or 1
ts = sum(l[4] for l in lines) or 1
print(f"kernel: {first}\nwarp instructions {ti}; samples {ts}")
print(f"{'file:line':24s} {'instr%':>7s} {'smpl%':>7s}  top stalls | source")
for f, ln, src, ie, sm, st in lines:
    if 100.0 * ie / ti < minpct and 100.0 * sm / ts < minpct:
        continue
    top = ", ".join(f"{k} {100*v//max(sm,1)}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{f+':'+str(ln):24s} {100.0*ie/ti:7.2f} {100.0*sm/ts:7.2f}  {top:40s} | {src[:90]}")
