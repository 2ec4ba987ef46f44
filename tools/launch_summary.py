"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import collections
import csv
import sys


def main(path, skip=0):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    gi = hdr.index('Grid Size') if 'Grid Size' in hdr else None
    agg = collections.OrderedDict()
    for r in rows[1 + skip:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        u = r[ui]
        v = v / 1e6 if u == 'ns' else v / 1e3 if u in ('us', 'usecond') else v * 1e3 if u == 's' else v
        name = r[ki]
        name = name.replace('<unnamed>::', '').replace('void ', '')
        if 'gemm_kernel' in name:
            name = name.split('(gemm::Operands')[0]
        else:
            name = name.split('(')[0]
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1; a[1] += v; a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    print(f'{"kernel":78s} {"launches":>8s} {"total ms":>10s} {"share":>7s} {"max ms":>9s}')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{k[:78]:78s} {a[0]:8d} {a[1]:10.3f} {100 * a[1] / tot:6.1f}% {a[2]:9.3f}')
    print(f'{"TOTAL":78s} {sum(a[0] for a in agg.values()):8d} {tot:10.3f}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
