#!/usr/bin/env python
"""Counts the Blackwell-specific SASS instructions per kernel in lib/*.o (cuobjdump -sass): UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, SHFL, HMMA (legacy
tensor path: must be absent).  No GPU needed.  Writes a table to stdout (committed as profiles/r02_sass_counts.txt)."""
import collections
import glob
import os
import re
import subprocess

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(HERE, "quaternion-convolutional-neural-networks-for-end-to-end-automatic-speech-recognition_b200", "lib")
PAT = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "SHFL", "HMMA",
       "RED", "ATOM"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    except Exception:
        return name


def short(name):
    d = demangle(name)
    d = re.sub(r"qnn::\(anonymous namespace\)::", "", d)
    d = re.sub(r"\(.*$", "", d)
    return d.replace("void ", "")


rows = []
for obj in sorted(glob.glob(os.path.join(LIB, "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn, counts = None, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if fn:
                rows.append((os.path.basename(obj), fn, counts))
            fn, counts = m.group(1), collections.Counter()
            continue
        if fn:
            m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                op = m.group(1).split(".")[0]
                counts["total"] += 1
                for p in PAT:
                    if op.startswith(p):
                        counts[p] += 1
    if fn:
        rows.append((os.path.basename(obj), fn, counts))

print("# SASS instruction counts per kernel (cuobjdump -sass lib/*.o, sm_100a); UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,")
print("# UTMALDG/UTMASTG = TMA load/store, UBLKCP = bulk copy; HMMA (legacy mma.sync path) must be 0 everywhere.")
cols = ["total"] + PAT
print("%-22s %-70s " % ("object", "kernel") + " ".join("%8s" % c for c in cols))
for obj, fn, c in rows:
    print("%-22s %-70s " % (obj, short(fn)[:70]) + " ".join("%8d" % c[k] for k in cols))
