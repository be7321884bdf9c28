"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line: samples, instructions, top stall."""
import csv
import sys


def main(path, kernel_filter='', top=40):
    rows = list(csv.reader(open(path)))
    cur_file, cur_fn, hdr = None, None, None
    agg = {}
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]
        elif r[0] == 'Function Name':
            cur_fn = r[1]
        elif r[0] == 'Line No':
            hdr = r
        elif hdr and len(r) == len(hdr) and r[2] == '-' and r[0].isdigit():     # per-source-line summary row
            if kernel_filter not in cur_fn:
                continue
            d = dict(zip(hdr, r))
            si = hdr.index('# Samples')
            stalls = {h: float(v or 0) for h, v in zip(hdr, r) if h.startswith('stall_') and 'Not Issued' not in h}
            key = (cur_fn[:50], cur_file, int(r[0]), r[1].strip()[:110])
            a = agg.setdefault(key, [0.0, 0.0, {}])
            a[0] += float(r[si] or 0)
            a[1] += float(d['Instructions Executed'] or 0)
            for k, v in stalls.items():
                a[2][k] = a[2].get(k, 0) + v
    tot = sum(a[0] for a in agg.values())
    toti = sum(a[1] for a in agg.values())
    print(f'total samples {tot:.0f}, warp instructions {toti:.3e}')
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        st = sorted(a[2].items(), key=lambda kv: -kv[1])[:2]
        sts = ' '.join(f'{k[6:]}={v / max(a[0], 1) * 100:.0f}%' for k, v in st)
        print(f'{a[0] / tot * 100:5.1f}% smp {a[1] / toti * 100:5.1f}% ins  {key[1]}:{key[2]:<4d} [{sts}]  {key[3]}')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '', int(sys.argv[3]) if len(sys.argv) > 3 else 40)
