"""
Write the judged summary of an ncu report into profiles/ (text, committed).

    python tools/summarize_profile.py gpurun_out/prof_c2.ncu-rep profiles/r01_c2_tensor.txt PAIRS "title"
"""
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


def main(rep, out, pairs, title):
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


if __name__ == '__main__':
    main(*sys.argv[1:5])
