"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
per kernel template: launches, total device time, share, DRAM bytes read / written.

    python tools/launch_summary.py launches.csv [> profiles/rNN_full_step_launches.txt]
"""
import collections
import csv
import sys


def to_ms(v, u):
    return v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v * 1e3 if u in ('s', 'second') else v


def to_mb(v, u):
    return v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1e-6)


def main(path):
    rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) > 5]
    hdr = next(r for r in rows if 'Kernel Name' in r)
    ki, mi, vi, ui, ii = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit', 'ID'))
    agg, ids = collections.OrderedDict(), set()
    for r in rows[rows.index(hdr) + 1:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        name = r[ki].replace('<unnamed>::', '').replace('void ', '')
        name = name.split('(gemm::Operands')[0] if ('gemm_kernel' in name or 'fuse_kernel' in name or 'ares_kernel' in name) else name.split('(')[0]
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        if r[mi] == 'gpu__time_duration.sum':
            a[0] += 1; a[1] += to_ms(v, r[ui]); ids.add(r[ii])
        elif r[mi] == 'dram__bytes_read.sum':
            a[2] += to_mb(v, r[ui])
        elif r[mi] == 'dram__bytes_write.sum':
            a[3] += to_mb(v, r[ui])
    tot = sum(a[1] for a in agg.values())
    print(f'{"kernel":84s} {"launches":>8s} {"total ms":>9s} {"share":>6s} {"DRAM rd MB":>10s} {"DRAM wr MB":>10s}')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{k[:84]:84s} {a[0]:8d} {a[1]:9.3f} {100 * a[1] / tot:5.1f}% {a[2]:10.1f} {a[3]:10.1f}')
    print(f'{"TOTAL":84s} {sum(a[0] for a in agg.values()):8d} {tot:9.3f} {"":6s} {sum(a[2] for a in agg.values()):10.1f} '
          f'{sum(a[3] for a in agg.values()):10.1f}')


if __name__ == '__main__':
    main(sys.argv[1])
