#!/usr/bin/env python
"""Per-kernel SASS opcode evidence of the shipped library: counts of the Blackwell-native mnemonics
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk, UTMALDG/UTMASTG = tensor-map TMA, SYNCS = mbarrier,
UTCBAR = tcgen05.commit) and of the legacy tensor path (HMMA) per kernel.  Runs without a GPU:
  python tools/sass_histogram.py > profiles/r2_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "factored-neus_b200", "libfneus_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "MUFU", "F2FP", "HADD2", "RED", "ATOM"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            for k in ("RED", "ATOM"):                      # REDG / ATOMG / ATOMS ...
                if op.startswith(k):
                    op = k
            per[cur][op] += 1
            per[cur]["__total__"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    print("# SASS opcode histogram of `factored-neus_b200/libfneus_b200.so` (sm_100a)\n")
    print("`cuobjdump -sass` of the shipped library, instructions counted per kernel (static counts). `UTCHMMA` = `tcgen05.mma`, "
          "`UTCBAR` = `tcgen05.commit`, `LDTM` = `tcgen05.ld`, `UBLKCP` = `cp.async.bulk` (1-D bulk copies of pre-swizzled operand "
          "images: no tensor maps, hence no `UTMALDG`), `SYNCS` = mbarrier operations; `HMMA` (legacy `mma.sync`) must be 0.\n")
    print("| kernel | total | " + " | ".join(KEYS) + " |")
    print("|---|---:|" + "---:|" * len(KEYS))
    tot = collections.Counter()
    for (name, c), dn in zip(per.items(), demangle):
        short = re.sub(r"\(.*", "", dn).replace("fneus::", "")
        if not any(c[k] for k in ("UTCHMMA", "LDTM", "UBLKCP", "HMMA")):
            continue
        print("| `%s` | %d | " % (short, c["__total__"]) + " | ".join(str(c[k]) for k in KEYS) + " |")
        tot.update(c)
    print("| **all tensor-core kernels** | %d | " % tot["__total__"] + " | ".join(str(tot[k]) for k in KEYS) + " |")
    n_other = sum(1 for c in per.values() if not any(c[k] for k in ("UTCHMMA", "LDTM", "UBLKCP", "HMMA")))
    print("\n%d further kernels (FP32 CUDA-core GEMMs, sampling, compositing, packing, loss, marching cubes) use none of the "
          "above tensor / bulk-copy instructions." % n_other)


if __name__ == "__main__":
    main()
