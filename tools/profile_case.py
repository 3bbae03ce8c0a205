"""
Run one benchmark workload a few times through the C ABI (host buffers), for ncu.

    python tools/profile_case.py c2 [reps] [--exact]

Same workloads as bench.py; no torch import, so ncu replays stay short.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import bench         # noqa: E402
from fatiando_b200 import prism   # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 3
    exact = "--exact" in sys.argv
    w = bench.make_workload(name, 1)
    if "--names" in sys.argv:                    # another component set on the same model and points
        w['names'] = sys.argv[sys.argv.index("--names") + 1].split(",")
    if "--points" in sys.argv:
        k = int(sys.argv[sys.argv.index("--points") + 1])
        for key in ("xp", "yp", "zp"):
            w[key] = np.ascontiguousarray(w[key][:k])
    with prism.Model(w['model'], w['prop'], keep_all=w['prop'] is None, fvec=w['fvec'],
                     node_sharing="--per-cell" not in sys.argv) as model:
        for _ in range(reps):
            t0 = time.perf_counter()
            if w['kind'].startswith('jacobian'):
                # device-resident row block, ONE launch (what bench.py times)
                from fatiando_b200 import normal
                rows = min(w['xp'].size, int(sys.argv[sys.argv.index('--rows') + 1]) if '--rows' in sys.argv else 8192)
                grid = w['kind'] == 'jacobian_grid'
                sel = slice(None) if grid else slice(0, rows)
                with normal.SensitivityBlock(model, w['xp'][sel], w['yp'][sel], w['zp'][sel], w['names'][0],
                                             per_cell="--per-cell" in sys.argv, grid_reuse=grid) as block:
                    pairs = block.shape[0] * model.size
                    ms, nl = block.build_ms, 1
                print("%s: %d entries, kernels %.3f ms -> %.3e entries/s (path %s)" % (
                    name, pairs, ms, pairs / (ms * 1e-3), block.path))
                continue
            else:
                model.fields(w['xp'], w['yp'], w['zp'], w['names'], fvec=w['fvec'],
                             reference_order=exact)
                pairs = w['xp'].size * model.size
            ms, nl = model.last_kernel_ms()
            print("%s: %d pairs, kernels %.3f ms (%d launches) -> %.3e pairs/s; call %.1f ms" % (
                name, pairs, ms, nl, pairs / (ms * 1e-3), 1e3 * (time.perf_counter() - t0)))


if __name__ == "__main__":
    main()
