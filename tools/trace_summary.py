#!/usr/bin/env python
"""Summary of a LETKF_EXP_TRACE dump (tag, clock64 pairs of CTA 0 / thread 0): per tag-to-tag transition the count,
median and mean clocks and the share of the traced time; and the mean clocks per point of each transition."""
import collections
import sys


def main(path):
    rows = [tuple(map(int, l.split())) for l in open(path) if l.strip()]
    d = collections.defaultdict(list)
    for (t0, c0), (t1, c1) in zip(rows, rows[1:]):
        if c1 >= c0:
            d[(t0, t1)].append(c1 - c0)
    npts = sum(1 for t, _ in rows if t == 2)
    tot = sum(sum(v) for v in d.values())
    print(path, "entries", len(rows), "points", npts, "clocks/point %.0f" % (tot / max(npts, 1)))
    for k, v in sorted(d.items()):
        v2 = sorted(v)
        print("%-10s n %5d  med %7d  mean %7d  per-point %8.0f  share %.3f" % (k, len(v), v2[len(v2) // 2], sum(v) // len(v),
                                                                         sum(v) / max(npts, 1), sum(v) / tot))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
