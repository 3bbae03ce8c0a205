"""
Executed warp instructions per SOURCE line of a kernel (ncu --set full --import-source on).

    python tools/ncu_lines.py report.ncu-rep [pairs] [top]
"""
import csv
import subprocess
import sys


def main(path, pairs=None, top=40):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass,cuda'],
                         capture_output=True, text=True).stdout.splitlines()
    rows, cur, hdr = [], None, None
    for l in out:
        if l.startswith('"File Path"'):
            cur = next(csv.reader([l]))[1].split('/')[-1]
        elif l.startswith('"Line No"'):
            hdr = next(csv.reader([l]))
        elif hdr and cur:
            r = next(csv.reader([l]))
            if len(r) < 8 or not r[0]:
                continue
            try:
                n = int(r[hdr.index('Instructions Executed')])
            except ValueError:
                continue
            if n:
                rows.append((n, cur, r[0], r[1].strip()[:100]))
    scale = 32.0 / float(pairs) if pairs else 1.0
    byfile = {}
    for n, f, ln, src in rows:
        byfile[f] = byfile.get(f, 0) + n
    print('per file:', {k: round(v * scale, 1) for k, v in sorted(byfile.items(), key=lambda kv: -kv[1])})
    for n, f, ln, src in sorted(rows, reverse=True)[:int(top)]:
        print('%8.1f  %s:%s  %s' % (n * scale, f, ln, src))


if __name__ == '__main__':
    main(*sys.argv[1:])
