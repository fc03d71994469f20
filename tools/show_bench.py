#!/usr/bin/env python
"""Prints the headline numbers of bench.py JSON lines: show_bench.py file.json [...]"""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "ERR", e)
        continue
    r, i = d.get("roofline") or {}, d.get("issue_roofline") or {}
    print("%s: %s N=%s value %.4g ms/step %.5g kernel_ms %.5g hbm %.4f issue %s e2e %s parity %s" % (
        path, (d.get("config") or {}).get("workload", "?")[:40], d.get("n_gpus"), d["value"], d["ms_per_step"],
        r.get("kernel_ms") or 0, r.get("frac") or 0, i.get("frac"), (d.get("e2e") or {}).get("value"),
        (d.get("parity_spot_check") or {}).get("ok")))
    for k, v in (d.get("configs") or {}).items():
        print("   %-4s ms/step %.5f kernel_ms %.5f hbm %.4f issue %s qps %.4g parity %s e2e %s / %s" % (
            k, v["ms_per_step"], v["kernel_ms"], v["hbm_frac"], v["issue_frac"], v["queries_per_s"],
            (v["parity_spot_check"] or {}).get("ok"), (v.get("e2e") or {}).get("value"),
            (v.get("e2e_compact") or {}).get("value")))
