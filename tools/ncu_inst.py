#!/usr/bin/env python
"""Per-source-line executed warp instructions from an .ncu-rep, sorted by file:line (needs -lineinfo)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
sections, cur = [], None
for r in rows:
    if r and r[0] == "File Path": cur = {"file": r[1], "rows": []}; sections.append(cur)
    elif r and r[0] == "Line No": cur["hdr"] = r
    elif cur is not None and "hdr" in cur and len(r) == len(cur["hdr"]): cur["rows"].append(r)
tot = 0
for s in sections:
    h = s["hdr"]; iN = h.index("Instructions Executed"); iS = h.index("# Samples")
    for r in s["rows"]:
        if not r[0]: continue
        try: n = int(r[iN])
        except ValueError: continue
        if n == 0: continue
        tot += n
        print(f"{s['file'].split('/')[-1]}:{int(r[0]):4d} {n:10d} {int(r[iS] or 0):6d}  {r[1][:100]}")
print("total", tot)
