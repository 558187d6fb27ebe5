"""Opcode histogram per kernel of libshb200.so (cuobjdump -sass): the evidence that the hot kernels are Blackwell-native --
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier.

    python scripts/sass_histogram.py > profiles/r02_sass_opcodes.md
"""
import collections, os, re, subprocess, sys

lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "semantichuman_b200", "libshb200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "ELECT", "R2UR", "LDGSTS", "HMMA", "FFMA", "MUFU", "LDG", "STG", "LDS", "STS"]
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1).split(".")[0]] += 1
        hist[kern]["_total"] += 1
print("# SASS opcode counts per kernel of libshb200.so (sm_100a), `cuobjdump -sass`\n")
print("| kernel | instructions | " + " | ".join(KEY) + " |")
print("|---|---|" + "---|" * len(KEY))
for k, c in hist.items():
    if "shb::" not in k:
        continue
    print(f"| `{k.replace('void ', '')}` | {c['_total']} | " + " | ".join(str(c.get(o, 0)) for o in KEY) + " |")
