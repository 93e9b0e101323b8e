"""Per-source-line totals of an ncu report's source page: `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > f.csv;
python tools/ncu_lines.py f.csv [file-substring] [lo] [hi]` prints, for the lines of the file in [lo, hi], warp-stall samples
and executed warp instructions, plus the dominant stall reasons."""
import csv, sys
path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else "rc_sampler.cu"
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0; hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
top = int(sys.argv[5]) if len(sys.argv) > 5 else 0
cur = None; hdr = None; rows = []
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if cur and want in cur and r[0].isdigit():
        ln = int(r[0])
        if lo <= ln <= hi: rows.append((ln, r))
if not hdr: sys.exit("no header")
ix = {h: i for i, h in enumerate(hdr)}
iS = ix["# Samples"]; iI = ix["Instructions Executed"]
stalls = [(h, i) for h, i in ix.items() if h.startswith("stall_") and "Not Issued" not in h]
num = lambda x: int(x) if x.lstrip("-").isdigit() else 0
tot = sum(num(r[iS]) for _, r in rows)
print(f"lines {len(rows)}  samples {tot}")
agg = {}
for ln, r in rows:
    a = agg.setdefault(ln, [0, 0, {}, r[1]])
    a[0] += num(r[iS]); a[1] += num(r[iI])
    for h, i in stalls:
        v = num(r[i])
        if v: a[2][h] = a[2].get(h, 0) + v
items = sorted(agg.items(), key=(lambda kv: -kv[1][0]) if top else (lambda kv: kv[0]))
if top: items = items[:top]
for ln, (s, ins, st, src) in items:
    if s == 0 and ins == 0: continue
    dom = ", ".join(f"{k[6:]} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{ln:5d} {s:7d} {100.0 * s / max(tot, 1):5.1f}% {ins:9d}  {src.strip()[:90]:90s} | {dom}")
