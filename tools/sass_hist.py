"""SASS evidence that the hot path is Blackwell-native: per-kernel counts of the tcgen05 / TMA / TMEM opcodes in libsgk.so.
    python tools/sass_hist.py > profiles/r2_sass_hist.md        (cuobjdump runs without a GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "supervised-gan_b200", "libsgk.so")
KEYS = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "UBLKCP", "LDGSTS", "HMMA", "FFMA",
        "MUFU", "LDG", "STG", "LDS", "STS", "RED", "ATOM", "SHFL", "BAR")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1).split(".")[0]
        per[cur][op] += 1
        per[cur]["_total"] += 1
        if op == "UTMALDG":
            dims = re.search(r"UTMALDG\.(\dD)", line)
            if dims:
                per[cur]["UTMALDG." + dims.group(1)] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
print("# SASS opcode histogram of supervised-gan_b200/libsgk.so (cuobjdump -sass, sm_100a)\n")
print("Blackwell-only opcodes: `UTCHMMA` = tcgen05.mma, `UTMALDG` = TMA tensor load (cp.async.bulk.tensor), `UTCBAR` = tcgen05.commit,")
print("`LDTM` = tcgen05.ld (TMEM -> registers), `SYNCS` = mbarrier ops.  `HMMA` (mma.sync) must be absent.\n")
print("| kernel | instructions | " + " | ".join(KEYS) + " |")
print("|---|---:|" + "---:|" * len(KEYS))
for (name, c), dm in zip(per.items(), demangle):
    short = re.sub(r"\(.*", "", dm).replace("void ", "").replace("sgk::", "")
    for k in KEYS:
        tot[k] += c[k]
    tot["_total"] += c["_total"]
    if not any(c[k] for k in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR")) and c["_total"] < 600:
        continue
    print("| %s | %d | " % (short[:70], c["_total"]) + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
print("| **all %d kernels** | %d | " % (len(per), tot["_total"]) + " | ".join(str(tot[k]) for k in KEYS) + " |")
dims = collections.Counter()
for c in per.values():
    for k, v in c.items():
        if k.startswith("UTMALDG."):
            dims[k] += v
print("\nTMA load ranks: " + ", ".join("%s x %d" % kv for kv in sorted(dims.items())))
