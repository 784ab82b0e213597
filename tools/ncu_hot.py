#!/usr/bin/env python
"""Top stall sites of one kernel from an .ncu-rep source page. Usage: ncu_hot.py rep kernel_regex [n]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat, "--launch-skip", "0", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
i_src, i_s, i_ex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
idx = {c: h.index(c) for c in stalls}
body = [r for r in rows[hi + 1:] if len(r) > i_s and r[0] != "Address"]
tot = sum(int(r[i_s] or 0) for r in body)
agg = {c: sum(int(r[idx[c]] or 0) for r in body) for c in stalls}
print("samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for r in sorted(body, key=lambda r: -int(r[i_s] or 0))[:n]:
    st = {c[6:]: int(r[idx[c]] or 0) for c in stalls if int(r[idx[c]] or 0)}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(r[i_s].rjust(6), r[i_ex].rjust(9), r[i_src][:64].ljust(64), st)
