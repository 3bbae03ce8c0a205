"""
Write the judged summary of an ncu report into profiles/ (text, committed).

    python tools/summarize_profile.py gpurun_out/prof_c2.ncu-rep profiles/r01_c2_tensor.txt PAIRS "title" [KEY [HASHFILE]]

With KEY it also writes profiles/r02_kernel_KEY.json, the numbers bench.py's roofline reads
(executed FP64-pipe instructions per pair, DRAM bytes per launch, ...) together with the hash
of the DEVICE sources that kernel is built from (bench.kernel_sources: its translation unit
and the headers it includes) and, separately, of the launch code (bench.launch_sources);
HASHFILE = JSON {key: hash, "launch": hash} the GPU box computed from its snapshot, default:
this tree's.  bench.py refuses the file when the kernel hash differs from its build's and
says so in its line when only the launch code changed.
"""
import json
import os
import re
import csv
import io
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__occupancy_limit_registers',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_write.sum.per_second',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__cycles_elapsed.avg.per_second',
]


def main(rep, out, pairs, title, key=None, hashfile=None, jsondir=None):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    name = vals[hdr.index('Kernel Name')]
    lines = ['# %s' % title, '# source: %s  (ncu --set full --clock-control none, one launch)' % rep,
             'kernel: %s' % name, '']
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append('%-72s %s %s' % (k, vals[i], units[i]))
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio'):
            lines.append('%-72s %s' % (h.replace('smsp__average_warps_issue_stalled_', 'stall/issue: '), vals[i]))
    mix = subprocess.run([sys.executable, __file__.replace('summarize_profile', 'ncu_hot'), rep, pairs],
                         capture_output=True, text=True).stdout
    lines += ['', '# dynamic instruction mix (ncu source page, executed warp instructions)', mix]
    open(out, 'w').write('\n'.join(lines))
    print('wrote', out)
    if key:
        def val(k):
            return float(vals[hdr.index(k)].replace(',', '')) if k in hdr else None

        def byts(k):
            if k not in hdr:
                return None
            mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}[units[hdr.index(k)]]
            return val(k) * mult
        m = re.search(r'per pair: total ([0-9.]+)\s+fp64 ([0-9.]+)', mix)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, root)
        import bench
        # HASHFILE: JSON {key: kernel hash} written by the GPU box from its snapshot
        if hashfile:
            hashes = json.load(open(hashfile))
            h, lh = hashes[key], hashes["launch"]
        else:
            h, lh = bench.kernel_hash(key), bench.launch_hash()
        d = {"key": key, "kernel": name, "pairs_per_launch": float(pairs),
             "inst_per_pair": float(m.group(1)), "fp64_inst_per_pair": float(m.group(2)),
             "time_ms": val('gpu__time_duration.sum'),
             "pipe_fp64_pct": val('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed'),
             "issue_active_pct": val('smsp__issue_active.avg.pct_of_peak_sustained_active'),
             "dram_bytes_per_launch": (byts('dram__bytes_read.sum') or 0) + (byts('dram__bytes_write.sum') or 0),
             "registers": val('launch__registers_per_thread'),
             "kernel_hash": h,
             "kernel_sources": [os.path.basename(f) for f in bench.kernel_sources(key)],
             "launch_hash": lh,
             "launch_sources": [os.path.basename(f) for f in bench.launch_sources()],
             "source": os.path.basename(rep) + " (summarised on the GPU box; the report itself is not kept)",
             "how": "ncu --set full --clock-control none --import-source on, one launch; instruction "
                    "counts from the source page (executed warp instructions x 32 / pairs)"}
        path = os.path.join(jsondir or os.path.join(root, 'profiles'), 'r02_kernel_%s.json' % key)
        json.dump(d, open(path, 'w'), indent=1)
        print('wrote', path)


if __name__ == '__main__':
    main(*sys.argv[1:8])
