#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[start]
    kn, mv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= mv:
            continue
        try:
            agg.setdefault(r[kn], []).append(float(r[mv].replace(',', '')))
        except ValueError:
            continue
    total = sum(sum(v) for v in agg.values())
    print('%-78s %5s %10s %10s %6s' % ('kernel', 'n', 'mean_us', 'sum_us', 'share'))
    for k, v in agg.items():
        print('%-78s %5d %10.1f %10.1f %5.1f%%' % (k[:78], len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3,
                                                   100 * sum(v) / total))


if __name__ == '__main__':
    main(sys.argv[1])
