"""Static SASS instruction mix per kernel of a .o/.so/.cubin (cuobjdump -sass)."""
import collections, re, subprocess, sys

FP64 = ('DFMA', 'DADD', 'DMUL', 'DSETP', 'DMNMX', 'F2F', 'I2F', 'F2I')

def main(path, pattern=''):
    txt = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
    for f in re.split(r'\n\s*Function : ', txt)[1:]:
        name = f.split('\n')[0]
        if pattern and pattern not in name:
            continue
        ops = collections.Counter()
        for line in f.split('\n'):
            m = re.search(r'/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
            if m:
                ops[m.group(1).split('.')[0]] += 1
        tot = sum(ops.values())
        d = sum(v for k, v in ops.items() if k in ('DFMA', 'DADD', 'DMUL', 'DSETP'))
        print('%s\n   total %d  fp64-pipe %d (%.0f%%)  MUFU %d' % (name[:90], tot, d, 100.0 * d / max(tot, 1), ops['MUFU']))
        print('   ', ops.most_common(16))

if __name__ == '__main__':
    main(*sys.argv[1:])
