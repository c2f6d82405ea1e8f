"""Instruction histogram of the shipped library (cuobjdump -sass), per kernel: the tcgen05 / TMA / TMEM mnemonics that prove the
Blackwell path, plus the packed-fp32 and shuffle counts of the parser kernels.  python tools/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "hyperpose_b200", "libhyperpose_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCBAR.MULTICAST", "UTMALDG", "UTMALDG.4D.IM2COL", "UTMALDG.2D.MULTICAST", "UTMASTG", "LDTM", "UTCATOMSWS", "SYNCS",
         "FFMA2", "FADD2", "FMUL2", "FFMA", "SHFL", "VOTE", "LDS", "STS", "LDG", "STG", "HMMA", "BAR", "UCGABAR_ARV"]
per = collections.OrderedDict()
cur = None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m and cur:
        op = m.group(1)
        per[cur]["__total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                per[cur][w] += 1
        if op.startswith("UTMALDG") and "IM2COL" in op:
            per[cur]["UTMALDG.4D.IM2COL"] += 1
        if op.startswith("UTMALDG") and "MULTICAST" in op:
            per[cur]["UTMALDG.2D.MULTICAST"] += 1
        if op.startswith("UTCBAR") and "MULTICAST" in op:
            per[cur]["UTCBAR.MULTICAST"] += 1
print(f"# cuobjdump -sass {os.path.basename(lib)}: instruction counts per kernel (sm_100a)")
tot = collections.Counter()
for k, c in per.items():
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    name = name.replace("(anonymous namespace)::", "").replace("hpb::", "")
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\((?!anonymous).*", "", name)[-90:]   # drop the parameter list, keep template arguments
    cols = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
    print(f"{name:<92} total={c['__total']:<6} {cols}")
    tot.update(c)
print("\n# whole library")
print(" ".join(f"{w}={tot[w]}" for w in WATCH if tot[w]), f"total={tot['__total']}")
print("# library dependencies (ldd): no cuBLAS / cuDNN / TensorRT / NCCL")
print(subprocess.run(["ldd", lib], capture_output=True, text=True).stdout)
