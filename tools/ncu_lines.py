#!/usr/bin/env python
"""Per-source-line executed instructions and stall samples from an .ncu-rep (needs -lineinfo)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
by_inst = len(sys.argv) > 3 and sys.argv[3] == "inst"   # sort by executed instructions instead of stall samples
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
sections, cur = [], None
for r in rows:
    if r and r[0] == "File Path": cur = {"file": r[1], "rows": []}; sections.append(cur)
    elif r and r[0] == "Line No": cur["hdr"] = r
    elif cur is not None and "hdr" in cur and len(r) == len(cur["hdr"]): cur["rows"].append(r)
agg = []
stall_cols = None
for s in sections:
    h = s["hdr"]; iN = h.index("Instructions Executed"); iS = h.index("# Samples")
    if stall_cols is None:
        stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    for r in s["rows"]:
        if not r[0]: continue
        try: n = int(r[iN])
        except ValueError: continue
        st = {c: int(r[i] or 0) for i, c in stall_cols}
        agg.append((int(r[iS] or 0), n, s["file"].split("/")[-1], r[0], r[1][:80], st))
tot_s = sum(a[0] for a in agg); tot_n = sum(a[1] for a in agg)
print(f"total samples {tot_s}, total warp instructions {tot_n}")
tot_st = {}
for a in agg:
    for k, v in a[5].items(): tot_st[k] = tot_st.get(k, 0) + v
print("stall totals:", {k: v for k, v in sorted(tot_st.items(), key=lambda kv: -kv[1]) if v})
for smp, n, f, ln, src, st in sorted(agg, key=lambda a: -(a[1] if by_inst else a[0]))[:top]:
    main = max(st.items(), key=lambda kv: kv[1])[0] if any(st.values()) else ""
    print(f"{smp:6d} {n:9d} {main:18s} {f}:{ln}: {src}")
