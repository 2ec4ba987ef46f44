"""ncu launch list (csv, see tools/gpu_profile.sh) -> profiles/rNN_full_traffic.json: DRAM bytes per launch of every GEMM kernel
template of one eager forward, keyed the way bench.py names the templates (`gemm_kernel<256>` ...): the `roofline.traffic` source.

    python tools/traffic_json.py gpurun_out/r02_full_step_launches.csv > profiles/r02_full_traffic.json
"""
import collections
import csv
import json
import re
import sys


def main(path):
    rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) > 5]
    hdr = next(r for r in rows if 'Kernel Name' in r)
    ki, mi, vi, ui = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit'))
    agg = collections.OrderedDict()
    for r in rows[rows.index(hdr) + 1:]:
        m = re.search(r'(gemm_\w*kernel)<(\d+)', r[ki])
        if not m:
            continue
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        a = agg.setdefault(f'{m.group(1)}<{m.group(2)}>', dict(launches=0, dram_read_bytes=0.0, dram_write_bytes=0.0, ncu_ms=0.0))
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(r[ui], 1.0)
        if r[mi] == 'gpu__time_duration.sum':
            a['launches'] += 1
            a['ncu_ms'] += v / 1e6 if r[ui] in ('ns', 'nsecond') else v / 1e3 if r[ui] in ('us', 'usecond') else v
        elif r[mi] == 'dram__bytes_read.sum':
            a['dram_read_bytes'] += v * scale
        elif r[mi] == 'dram__bytes_write.sum':
            a['dram_write_bytes'] += v * scale
    for a in agg.values():
        a['traffic_per_launch_bytes'] = (a['dram_read_bytes'] + a['dram_write_bytes']) / max(a['launches'], 1)
    print(json.dumps({'source': 'ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum '
                                '--clock-control none python bench.py --profile-step (exactly one eager whole forward, round-2 kernels, default '
                                'precision plan), B200', 'per_step': agg}, indent=1))


if __name__ == '__main__':
    main(sys.argv[1])
