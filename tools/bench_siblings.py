"""
Throughput of the sibling families (spheres, polygonal prisms) through the Python API with
host buffers.  python tools/bench_siblings.py [sphere|polyprism [NBODIES]]
(NBODIES: a smaller model for an ncu capture -- the per-pair counts do not depend on it)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fatiando_b200 import polyprism, sphere, gridder   # noqa: E402
from fatiando_b200.mesher import PolygonalPrism, Sphere   # noqa: E402


def main():
    rs = np.random.RandomState(0)
    nbodies = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    t = np.linspace(0, 2 * np.pi, 20, endpoint=False)[::-1]
    polys = [PolygonalPrism(np.transpose([cx + 80 * np.cos(t), cy + 60 * np.sin(t)]), 50, 400,
                            {"density": 500., "magnetization": [1., 2., 3.]})
             for cx, cy in rs.uniform(0, 1e4, (nbodies or 5000, 2))]
    spheres = [Sphere(rs.uniform(0, 1e4), rs.uniform(0, 1e4), rs.uniform(100, 2000), 20.,
                      {"density": 500., "magnetization": [1., 2., 3.]}) for _ in range(nbodies or 100000)]
    x, y, z = gridder.regular((0, 1e4, 0, 1e4), (400, 400), z=-50)
    sets = ((["gz"], "density"), (["gxx", "gxy", "gxz", "gyy", "gyz", "gzz"], "density"),
            (["tf", "bx", "by", "bz"], "magnetization"))
    only = sys.argv[1] if len(sys.argv) > 1 else ""         # "sphere" / "polyprism": one family (for ncu)
    for label, module, model, unit in (("polyprism (20-vertex prisms)", polyprism, polys, 20),
                                       ("sphere", sphere, spheres, 1)):
        if only and not label.startswith(only):
            continue
        for names, prop in sets:
            with module.Model(model, prop) as mod:
                for _ in range(3):
                    t0 = time.perf_counter()
                    mod.fields(x, y, z, names, fvec=[0.5, 0.5, 0.7071])
                    dt = time.perf_counter() - t0
            print("%-28s %-34s %8.1f ms  %.3e rows x points/s" % (label, "+".join(names), dt * 1e3,
                                                                 unit * len(model) * x.size / dt))


if __name__ == "__main__":
    main()
