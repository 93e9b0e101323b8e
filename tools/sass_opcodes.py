#!/usr/bin/env python
"""Blackwell evidence per kernel: counts of the SASS mnemonics that prove the tcgen05 / TMEM / TMA / DMMA paths, plus
registers and spills from the ptxas logs.  Regenerates profiles/r02_sass_opcodes.txt from the built objects:
    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import glob, os, re, subprocess, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "redclust.jl_b200", "csrc")
WATCH = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "DMMA", "UBLKCP", "UBLKPF", "UTMALDG", "UTMASTG", "LDGSTS", "SYNCS",
         "REDUX", "BAR.SYNC", "ATOMS", "DADD", "DMUL", "DFMA", "IMMA", "HMMA"]


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n


def short(name):
    d = demangle(name)
    d = re.sub(r"\(anonymous namespace\)::", "", d)
    d = re.sub(r"_GLOBAL__N__[0-9a-f_]+rc_\w+_cu_[0-9a-f]+::", "", d)
    return d.split("(")[0][:90]


print("# SASS mnemonics per kernel (cuobjdump -sass of redclust.jl_b200/csrc/*.o, sm_100a) and ptxas resource usage")
print("# mnemonic legend: UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UBLKCP = cp.async.bulk (TMA engine, 1-D),")
print("#   UBLKPF = cp.async.bulk.prefetch.L2, UTMALDG = cp.async.bulk.tensor, LDGSTS = cp.async, SYNCS = mbarrier, DMMA = fp64 mma.sync,")
print("#   REDUX = warp-wide integer reduction")
for obj in sorted(glob.glob(os.path.join(CSRC, "*.o"))):
    if obj.endswith("_stats.o"):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    res = {}
    log = obj + ".ptxas.log"
    if os.path.exists(log):
        cur = None
        for line in open(log):
            m = re.search(r"Function properties for (\S+)", line) or re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                cur = m.group(1)
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and cur:
                res.setdefault(cur, {})["spill"] = (int(m.group(2)), int(m.group(3)))
            m = re.search(r"Used (\d+) registers", line)
            if m and cur:
                res.setdefault(cur, {})["regs"] = int(m.group(1))
    print(f"\n## {os.path.basename(obj)}")
    fn, counts, total = None, None, 0
    out = []
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            if fn:
                out.append((fn, counts, total))
            fn, counts, total = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if m and fn:
            op = m.group(1)
            total += 1
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w.endswith(".SYNC") and op.startswith(w)):
                    counts[w] += 1
    if fn:
        out.append((fn, counts, total))
    for fn, counts, total in sorted(out, key=lambda t: -t[2]):
        r = res.get(fn, {})
        tags = ", ".join(f"{k} {v}" for k, v in sorted(counts.items()) if v)
        sp = r.get("spill", (0, 0))
        print(f"{short(fn):60s} {total:6d} instr, {r.get('regs', '?'):>3} regs, spills {sp[0]}/{sp[1]} B | {tags}")
