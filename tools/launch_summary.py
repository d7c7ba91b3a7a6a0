#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals and the launch sequence
of the last complete training step.  Usage: python tools/launch_summary.py gpurun_out/launches.csv [--seq]"""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, seq = None, []
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        seq.append((d["Kernel Name"], v, d.get("Grid Size")))
    return seq


def main():
    seq = load(sys.argv[1])
    idx = [i for i, s in enumerate(seq) if "upsample_step" in s[0] or "upsample_iter" in s[0]]
    start, end = idx[-8], idx[-4]          # 4 up-sampling launches per step: the last complete step
    step = seq[start:end]
    tot = sum(v for _, v, _ in step)
    agg = collections.OrderedDict()
    for k, v, _ in step:
        a = agg.setdefault(k[:90], [0, 0.0])
        a[0] += 1
        a[1] += v
    print("one training step: %d launches, %.1f us (cold-cache, serialised)" % (len(step), tot))
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:24]:
        print("| `%s` | %d | %.1f | %.1f%% |" % (k, a[0], a[1], 100 * a[1] / tot))
    if "--seq" in sys.argv:
        for k, v, g in step:
            if "fneus" in k:
                print("%-70s %8.1f %s" % (k[:70], v, g))


if __name__ == "__main__":
    main()
