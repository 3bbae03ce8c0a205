"""
Cost of the per-point fallback and rate of a host-resident Jacobian (item 8).

  1. C5 (1e6-prism relief, 15625 points, gz+tensor): the clean grid against the same grid with
     ONE point moved into a hazard wall -- kernel time and which values change.
  2. C4 mesh: rows of the gz sensitivity matrix written into a HOST array, GB/s end to end.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np          # noqa: E402
import bench                # noqa: E402
from fatiando_b200 import prism   # noqa: E402
from fatiando_b200.flatten import flatten   # noqa: E402


def main():
    w = bench.make_workload("c5t", 1)
    flat = flatten(w['model'], 'density')
    hz = [h for h in flat.surface.hazard]
    print("hazard walls per axis:", [h.size for h in hz])
    xp, yp, zp = w['xp'].copy(), w['yp'].copy(), w['zp'].copy()
    with prism.Model(flat, 'density') as model:
        for _ in range(2):
            clean = model.fields(xp, yp, zp, w['names'])
        t_clean = model.last_kernel_ms()[0]
        assert model.last_fallback_points() == 0
        xp2 = xp.copy()
        xp2[1234] = hz[0][hz[0].size // 2]
        for _ in range(2):
            hit = model.fields(xp2, yp, zp, w['names'])
        t_hit = model.last_kernel_ms()[0]
        k = model.last_fallback_points()
        model.fields(xp2, yp, zp, w['names'], per_cell=True)
        t_cells = model.last_kernel_ms()[0]
        cells = model.fields(xp2[1234:1235].copy(), yp[1234:1235].copy(), zp[1234:1235].copy(), w['names'],
                             per_cell=True)
        others = np.arange(xp.size) != 1234
        same_others = all(np.array_equal(hit[f][others], clean[f][others]) for f in hit)
        same_point = all(hit[f][1234] == cells[f][0] for f in hit)
        print("C5 gz+tensor, 15625 points: clean %.2f ms; one point in a hazard wall %.2f ms (+%.2f %%), "
              "%d point(s) re-evaluated; whole call per cell %.2f ms" % (t_clean, t_hit,
              100 * (t_hit / t_clean - 1), k, t_cells))
        print("   flagged point bit-identical to a per-cell call on that point: %s; other points unchanged: %s"
              % (same_point, same_others))
    w = bench.make_workload("c4full", 1)
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    xr, yr, zr = (np.ascontiguousarray(w[k][:rows]) for k in ('xp', 'yp', 'zp'))
    from fatiando_b200 import pinned_empty
    with prism.Model(w['model'], None, keep_all=True) as model:
        shape = (rows, model.size)
        cases = [("fresh numpy.zeros per call", lambda: None),
                 ("the same pageable array again", None),
                 ("page-locked array (fatiando_b200.pinned_empty)", None)]
        reuse = np.zeros(shape)
        reuse[:] = 1.0
        pinned = pinned_empty(shape)
        for label, out in (("fresh numpy.zeros per call", None), ("the same pageable array again", reuse),
                           ("page-locked array (fatiando_b200.pinned_empty)", pinned)):
            for rep in range(3):
                t0 = time.perf_counter()
                j = model.jacobian(xr, yr, zr, 'gz', out=out)
                dt = time.perf_counter() - t0
                print("host-resident J, %s: %d x %d (%.2f GB) in %.3f s = %.1f GB/s end to end (kernels %.1f ms)"
                      % (label, rows, j.shape[1], j.nbytes / 1e9, dt, j.nbytes / 1e9 / dt, model.last_kernel_ms()[0]))
            if out is None:
                del j
        assert np.array_equal(reuse, pinned)


if __name__ == "__main__":
    main()
