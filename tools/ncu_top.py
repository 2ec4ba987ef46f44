"""Top stall sites from `ncu -i rep --page source --csv --kernel-name regex:...` output."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    # find header rows (they start with "Address")
    out = []
    hdr = None
    for r in rows:
        if r and r[0] == 'Address':
            hdr = r
            si, src, ie = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
            continue
        if hdr is None or len(r) <= max(si, src, ie) or not r[0].startswith('0x'):
            continue
        try:
            out.append((int(r[si] or 0), int(r[ie] or 0), r[src].strip(), len(out)))
        except ValueError:
            pass
    tot = sum(d[0] for d in out) or 1
    print('total samples', tot, 'warp instructions', sum(d[1] for d in out), 'sass lines', len(out))
    for s, e, t, i in sorted(out, reverse=True)[:top]:
        print(f'{s:7d} {100 * s / tot:5.1f}%  exec={e:10d}  #{i:5d}: {t[:110]}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
