#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` export: stall reasons, hottest instructions, opcode histogram.
    python tools/sass_summary.py gpurun_out/xxx_sass.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; data = rows[hi + 1:]
ix = {h: i for i, h in enumerate(hdr)}
print(rows[0][1][:120] if len(rows[0]) > 1 else "")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}; total = 0; ops = {}; hot = []
for n, r in enumerate(data):
    try: s = int(r[ix["# Samples"]])
    except Exception: continue
    total += s
    for k in stalls: tot[k] += int(r[ix[k]] or 0)
    m = r[ix["Source"]].split()
    if not m: continue
    k = m[1] if m[0].startswith("@") and len(m) > 1 else m[0]
    d = ops.setdefault(k.split(".")[0], [0, 0]); d[0] += s; d[1] += int(r[ix["Instructions Executed"]] or 0)
    hot.append((s, n, r[ix["Source"]].strip()[:90], {k[6:]: int(r[ix[k]] or 0) for k in stalls if int(r[ix[k]] or 0) > 0}))
print("samples", total)
print("stalls:", ", ".join(f"{k[6:]} {v / max(total,1):.2f}" for k, v in sorted(tot.items(), key=lambda x: -x[1])[:10]))
print("opcodes (samples, executed):", ", ".join(f"{k} {v[0]}/{v[1]}" for k, v in sorted(ops.items(), key=lambda x: -x[1][0])[:16]))
for s, n, src, st in sorted(hot, reverse=True)[:top]:
    print(f"{s:6d} #{n:5d} {src:90s} {dict(sorted(st.items(), key=lambda x: -x[1])[:3])}")
