"""Dev tool: aggregate an `ncu --page source --print-source cuda,sass --csv` export by CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
agg = {}
for r in rows[hi + 1:]:
    if len(r) <= iI or not r[0].isdigit() or r[2] != "-":
        continue                      # keep the per-line summary rows (Address == "-")
    agg[int(r[0])] = (int(r[iS] or 0), int(r[iI] or 0), r[1])
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print("samples", ts, "warp-inst", ti)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for ln, (s, i, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*s/ts:5.1f}% samp {100*i/ti:5.1f}% inst  L{ln:<4} {src.strip()[:120]}")
