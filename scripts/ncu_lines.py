#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an ncu report + the .so it ran.
usage: ncu_lines.py report.ncu-rep lib.so kernel_substr [topN]"""
import collections, csv, io, re, subprocess, sys, tempfile, os
rep, lib, ksub = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the csv holds one table per kernel launch, each introduced by a "Kernel Name" row
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
blk = [b for b in blocks if ksub in b["name"]][0]
hdr = blk["rows"][0]
iA, iS, iE, iSm = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
addr2line = {}
mangled_hint = re.sub(r"[^A-Za-z0-9_]", "", ksub)
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    secs = re.split(r"\n//-+ \.text\.", txt)
    for sec in secs[1:]:
        name = sec.split(" ", 1)[0]
        demangled = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip()
        def norm(x):
            x = x.replace("(int)", "").replace("(bool)", "").replace("true", "1").replace("false", "0")
            x = x.replace("void ", "").replace(" ", "")
            return x[:x.index(">(") + 1] if ">(" in x else x.split("(")[0]
        if norm(blk["name"]) != norm(demangled):
            continue
        cf = cl = None
        for l in sec.split("\n"):
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m: cf, cl = m.group(1).split("/")[-1], int(m.group(2)); continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m: addr2line[int(m.group(1), 16)] = (cf, cl)
rows = blk["rows"][1:]
base = int(rows[0][iA], 16)
by, bys, tot, stot = collections.Counter(), collections.Counter(), 0, 0
for r in rows:
    if len(r) <= iE: continue
    k = addr2line.get(int(r[iA], 16) - base, ("?", 0))
    n = int(r[iE]); by[k] += n; tot += n; s = int(r[iSm]); bys[k] += s; stot += s
print(f"{blk['name'][:90]}\n total warp instr {tot}, samples {stot}, mapped lines {len(addr2line)}")
srcs = {}
order = sorted(by, key=lambda k: -bys[k]) if os.environ.get("BY_SAMPLES") else [k for k, _ in by.most_common()]
for (f, l) in order[:top]:
    n = by[(f, l)]
    text = ""
    for d in ("rfi_toolbox_b200/csrc", "include"):
        pth = os.path.join(d, f or "")
        if os.path.exists(pth):
            srcs.setdefault(pth, open(pth).read().split("\n"))
            if 0 < l <= len(srcs[pth]): text = srcs[pth][l - 1].strip()[:90]
    print(f"{100*n/tot:5.1f}% instr {100*bys[(f,l)]/max(stot,1):5.1f}% smp  {f}:{l}  {text}")
