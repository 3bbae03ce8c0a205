"""
Static evidence from the shipped library (no GPU needed): per kernel, registers and local
memory from `cuobjdump -res-usage`, and the SASS mnemonics that show what the kernel uses --
TMA bulk copies (UBLKCP), mbarrier waits (SYNCS), fp64 tensor instructions (DMMA), local-memory
traffic (STL / LDL: spills), FP64-pipe instructions.

    python tools/sass_evidence.py [library] > profiles/r02_sass_evidence.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.split('\n')
    return dict(zip(names, out))


def main(lib):
    res = subprocess.run(['cuobjdump', '-res-usage', lib], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r'Function (\S+):\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)', res):
        usage[m.group(1)] = tuple(int(v) for v in m.groups()[1:])
    sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    rows = []
    for f in re.split(r'\n\s*Function : ', sass)[1:]:
        name = f.split('\n')[0].strip()
        ops = collections.Counter()
        for line in f.split('\n'):
            m = re.search(r'/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
            if m:
                ops[m.group(1).split('.')[0]] += 1
        rows.append((name, ops))
    pretty = demangle([r[0] for r in rows])
    print('# %s  (cuobjdump -res-usage / -sass of the shipped sm_100a library; static counts)' % os.path.relpath(lib, ROOT))
    print('# %-92s %4s %5s %6s | %5s %5s %6s %5s %5s %6s' % ('kernel', 'regs', 'stack', 'smem', 'STL', 'LDL', 'UBLKCP', 'SYNCS', 'DMMA', 'FP64'))
    tot_spill = 0
    for name, ops in sorted(rows, key=lambda r: pretty[r[0]]):
        reg, stack, shared, local = usage.get(name, (0, 0, 0, 0))
        fp64 = sum(ops[k] for k in ('DFMA', 'DADD', 'DMUL', 'DSETP'))
        short = re.sub(r'\bfb200::', '', pretty[name])
        short = re.sub(r'\(unsigned int\)', '', short)
        short = re.sub(r'\((?:fb200::)?\w+(?: const)?\s*[*&]?\)$', '', short)
        short = re.sub(r'^void ', '', short)
        tot_spill += ops['STL'] + ops['LDL']
        print('  %-92s %4d %5d %6d | %5d %5d %6d %5d %5d %6d' % (short[:92], reg, stack, shared, ops['STL'], ops['LDL'],
                                                                   ops['UBLKCP'], ops['SYNCS'], ops['DMMA'], fp64))
    print('# kernels: %d; STL + LDL over all of them: %d' % (len(rows), tot_spill))


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'fatiando_b200', 'libfatiando_b200.so'))
