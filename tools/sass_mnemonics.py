"""Per-kernel histogram of the Blackwell-specific SASS mnemonics in the shipped library (cuobjdump -sass):
UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit), SYNCS (mbarrier), UTMALDG (tensor-map
TMA: not used -- operands are K8-blocked so a contiguous bulk copy lands in the canonical layout), HGMMA (sm_90 wgmma: must be 0).

    python tools/sass_mnemonics.py [> profiles/r02_sass_mnemonics.txt]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'gpemsr_b200', 'lib', 'libgpemsr_b200.so')
KEYS = ['UTCHMMA', 'LDTM', 'UBLKCP', 'UTCBAR', 'SYNCS', 'UTMALDG', 'HGMMA', 'HMMA', 'FFMA', 'MUFU.EX2', 'F2FP']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(['cu++filt', n], capture_output=True, text=True).stdout.strip() or n
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r'^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?PT?\d*\s+)?([A-Z0-9_.]+)', line)
        if m:
            op = m.group(1)
            cur['_total'] += 1
            for k in KEYS:
                if op == k or op.startswith(k + '.') or (k == 'MUFU.EX2' and op.startswith('MUFU.EX2')):
                    cur[k] += 1
    tot = collections.Counter()
    print(f'# {os.path.relpath(LIB, ROOT)}: {len(per)} kernels (sm_100a SASS); columns = instruction counts per kernel')
    print('# ' + ' '.join(f'{k:>8s}' for k in KEYS) + '    total  kernel')
    for name, c in sorted(per.items(), key=lambda kv: -kv[1]['UTCHMMA']):
        tot.update(c)
        short = re.sub(r'\(anonymous namespace\)::|\(int\)|\(bool\)', '', demangle(name))
        short = re.sub(r'\(.*', '', short).replace('void ', '')
        print('  ' + ' '.join(f'{c[k]:8d}' for k in KEYS) + f' {c["_total"]:8d}  {short[:150]}')
    print('# ' + ' '.join(f'{tot[k]:8d}' for k in KEYS) + f' {tot["_total"]:8d}  ALL')
    assert tot['HGMMA'] == 0 and tot['UTCHMMA'] > 0


if __name__ == '__main__':
    main()
