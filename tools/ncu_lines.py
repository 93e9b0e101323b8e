#!/usr/bin/env python
"""Aggregate the warp-stall samples of an ncu report by CUDA source line.
usage: tools/ncu_lines.py report.ncu-rep object.o kernel_substring [topN]
(ncu --page source gives SASS addresses; nvdisasm -g maps SASS offsets to file:line.)"""
import csv, io, re, subprocess, sys, tempfile, os, collections

rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
line_of = {}
infn, cur = False, None
for l in dis:
    if l.startswith(".text.") and l.endswith(":"):
        infn = kname in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
base = min(int(r[ix["Address"]], 16) for r in body)
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in body:
    off = int(r[ix["Address"]], 16) - base
    key = line_of.get(off, ("?", 0))
    n = int(r[ix["# Samples"]] or 0)
    agg[key]["samples"] += n
    agg[key]["inst"] += int(r[ix["Instructions Executed"]] or 0)
    tot["samples"] += n
    for s in stalls:
        v = int(r[ix[s]] or 0)
        agg[key][s] += v
        tot[s] += v
print("total samples", tot["samples"])
for s, v in tot.most_common(9):
    if s != "samples":
        print(f"  {s:28s} {100 * v / tot['samples']:5.1f}%")
srcs = {}
for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    f, ln = key
    text = ""
    for d in ("redclust.jl_b200/csrc", "."):
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", d, f)
        if os.path.exists(p):
            srcs.setdefault(p, open(p).read().splitlines())
            if 0 < ln <= len(srcs[p]):
                text = srcs[p][ln - 1].strip()[:90]
            break
    tops = ", ".join(f"{s[6:]}={100 * v / max(c['samples'], 1):.0f}%" for s, v in c.most_common(4) if s.startswith("stall_"))
    print(f"{100 * c['samples'] / tot['samples']:5.1f}%  {f}:{ln:<5d} inst={c['inst']:<12d} [{tops}]  {text}")
