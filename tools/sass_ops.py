"""Per-kernel counts of the SASS mnemonics that prove what the kernels use (tcgen05 MMA = UTCHMMA, TMA = UTMALDG /
UBLKCP, TMEM loads = LDTM, packed fp32 = FFMA2 / FADD2, 128-bit global loads) from the built library.
usage: python tools/sass_ops.py [path/to/libhnsw_b200.so] > profiles/r2_sass_ops.txt"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "pgvector-hnsw-partitioning_b200", "libhnsw_b200.so")
OPS = ["UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "SYNCS", "FFMA2", "FADD2", "FFMA", "LDG.E.NA.128", "LDG.E.128", "LDS.128", "SHFL", "VOTE", "ATOMS"]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
name, cnt, total = None, None, {}
rows = []
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        if name:
            rows.append((name, total[name], cnt))
        name = m.group(1)
        cnt = collections.Counter()
        total[name] = 0
        continue
    if name and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
        total[name] += 1
        for op in OPS:
            if re.search(r"\b" + re.escape(op) + r"(\b|\.)", line):
                cnt[op] += 1
                break
if name:
    rows.append((name, total[name], cnt))
demangled = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
print("# %s: %d kernels (sm_100a SASS)" % (os.path.basename(so), len(rows)))
print("# columns: SASS instructions, then counts of " + " ".join(OPS))
for (nm, tot, c), dn in sorted(zip(rows, demangled), key=lambda x: x[1]):
    dn = re.sub(r"\(.*", "", dn).replace("void ", "").replace("hb::", "")
    print("%-78s %6d  " % (dn[:78], tot) + " ".join("%s=%d" % (op, c[op]) for op in OPS if c[op]))
