"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        v *= {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(r[ui], 1.0)
        name = r[ki].split('(')[0].replace('b200tts::<unnamed>::', '').replace('void ', '')[:64]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f'{"kernel":66s} {"n":>6s} {"total_us":>12s} {"avg_us":>10s} {"share":>7s}')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{k:66s} {v[0]:6d} {v[1] / 1e3:12.1f} {v[1] / v[0] / 1e3:10.2f} {100 * v[1] / tot:6.1f}%')
    print(f'{"TOTAL":66s} {sum(v[0] for v in agg.values()):6d} {tot / 1e3:12.1f}')


if __name__ == '__main__':
    main(sys.argv[1])
