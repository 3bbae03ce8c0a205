"""
Dynamic instruction mix of a kernel from an ncu report (needs --set full --import-source on).

    python tools/ncu_hot.py report.ncu-rep [pairs]

Prints executed warp instructions by opcode and by pipe class, per prism x point
pair when `pairs` is given, and the stall-sample ranking of the hottest SASS lines.
"""
import collections
import csv
import io
import re
import subprocess
import sys

FP64 = {'DFMA', 'DADD', 'DMUL', 'DSETP', 'DMNMX'}
XU = {'MUFU', 'I2F', 'F2I', 'F2F', 'I2FP', 'F2FP'}


def main(path, pairs=None):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'],
                         capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    print(lines[start - 1][:150])
    rows = list(csv.DictReader(io.StringIO('\n'.join(lines[start:]))))
    ops = collections.Counter()
    samples = collections.Counter()
    total = 0
    for r in rows:
        m = re.match(r'\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', r['Source'])
        if not m:
            continue
        n = int(r['Instructions Executed'] or 0)
        ops[m.group(1)] += n
        total += n
        samples[m.group(1)] += int(r['# Samples'] or 0)
    warp_pairs = (float(pairs) / 32.0) if pairs else None
    fp64 = sum(v for k, v in ops.items() if k in FP64)
    print('executed warp instructions: %.4g   fp64-pipe: %.4g (%.1f%%)' % (total, fp64, 100.0 * fp64 / total))
    if warp_pairs:
        print('per pair: total %.1f   fp64 %.1f' % (total / warp_pairs, fp64 / warp_pairs))
    tot_samples = sum(samples.values())
    for k, v in ops.most_common(28):
        extra = '' if not warp_pairs else '  %7.1f /pair' % (v / warp_pairs)
        print('  %-8s %12d %5.1f%%%s   stall-samples %5.1f%%' % (k, v, 100.0 * v / total, extra,
                                                             100.0 * samples[k] / max(tot_samples, 1)))


if __name__ == '__main__':
    try:
        main(*sys.argv[1:])
    except BrokenPipeError:
        pass
