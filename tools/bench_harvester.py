"""
Planting inversion (fatiando_b200.harvester) on a cookbook-sized problem.

    python tools/bench_harvester.py [--cpu-accretions K]

Problem: cookbook/gravmag_harvester_tensor.py scaled like the reference's recipe --
gyy, gyz, gzz on a 51x51 grid (7803 data), PrismMesh 15x50x50 (37500 cells), 16 seeds in
two bodies.  Prints one JSON line: accretions, seconds and accretions/s of the CUDA engine;
with --cpu-accretions K also the same host logic on the numpy engine (tests/
harvest_numpy_engine.py: the reference's own per-candidate numpy reductions and one oracle
call per cell and data set) for the first K accretions, on one core like the reference.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from fatiando_b200 import harvester, prism, mesher, gridder   # noqa: E402


def problem():
    bounds = [0, 5000, 0, 5000, 0, 1500]
    P = mesher.Prism
    model = [P(500, 4500, 3000, 3500, 200, 700, {'density': 1200}),
             P(3000, 4500, 1800, 2300, 200, 700, {'density': 1200}),
             P(300, 2500, 1800, 2700, 500, 1000, {'density': -1000}),
             P(1800, 3700, 500, 1500, 300, 1300, {'density': -1000})]
    x, y, z = gridder.regular(bounds[:4], (51, 51), z=-150)
    rs = np.random.RandomState(0)
    f = prism.fields(x, y, z, model, ['gyy', 'gyz', 'gzz'])
    data = [harvester.Gyy(x, y, z, f['gyy'] + rs.normal(0, 2, x.size)),
            harvester.Gyz(x, y, z, f['gyz'] + rs.normal(0, 2, x.size)),
            harvester.Gzz(x, y, z, f['gzz'] + rs.normal(0, 2, x.size))]
    mesh = mesher.PrismMesh(bounds, (15, 50, 50))
    locs = [(xs, 3250, 600, {'density': 1200}) for xs in (800, 1200, 1700, 2100, 2500, 2900, 3300, 3700, 4200)]
    locs += [(xs, 2050, 600, {'density': 1200}) for xs in (3300, 3600, 4000, 4300)]
    locs += [(1000, 2200, 700, {'density': -1000}), (2000, 2200, 700, {'density': -1000}),
             (2700, 1000, 800, {'density': -1000})]
    return data, mesh, locs


def main():
    data, mesh, locs = problem()
    seeds = harvester.sow(locs, mesh)
    harvester.harvest(data, harvester.sow(locs[:1], mesh), mesh, 1.0, 0.0001)     # warm-up
    t0 = time.perf_counter()
    est, pred, rep = harvester.harvest(data, seeds, mesh, compactness=1.0, threshold=0.0001, report=True)
    dt = time.perf_counter() - t0
    line = {"what": "harvester.harvest, gyy+gyz+gzz 3x2601 data, PrismMesh 15x50x50, %d seeds" % len(seeds),
            "accretions": rep['accretions'], "seconds": dt, "accretions_per_s": rep['accretions'] / dt,
            "gpu_launches": rep['gpu_launches'], "misfit": rep['misfit'], "goal": rep['goal']}
    if "--cpu-accretions" in sys.argv:
        k = int(sys.argv[sys.argv.index("--cpu-accretions") + 1])
        from harvest_numpy_engine import NumpyEngine
        t0 = time.perf_counter()
        n = -1
        for n, up in enumerate(harvester.iharvest(data, seeds, mesh, 1.0, 0.0001, [],
                                                  engine_factory=NumpyEngine)):
            if n >= k:
                break
        dtc = time.perf_counter() - t0
        line["cpu_port"] = {"accretions": n, "seconds": dtc, "accretions_per_s": n / dtc, "cores": 1,
                            "kind": "port (reference numpy reductions + oracle effects, 1 core)"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
