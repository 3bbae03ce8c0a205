"""
3D inversion of gravity-gradient data by planting anomalous densities -- after the
reference's cookbook/gravmag_harvester_tensor.py, on the device-resident effect engine.

    python examples/harvester_tensor.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fatiando_b200 import gridder, harvester, mesher, prism   # noqa: E402
from tools.synthetic import noisy   # noqa: E402

bounds = [0, 5000, 0, 5000, 0, 1500]
model = [mesher.Prism(500, 4500, 3000, 3500, 200, 700, {'density': 1200}),
         mesher.Prism(3000, 4500, 1800, 2300, 200, 700, {'density': 1200})]
x, y, z = gridder.regular(bounds[:4], (51, 51), z=-150)
clean = prism.fields(x, y, z, model, ['gyy', 'gyz', 'gzz'])
gyy, gyz, gzz = (noisy(clean[f], 2, seed=k) for k, f in enumerate(['gyy', 'gyz', 'gzz']))

mesh = mesher.PrismMesh(bounds, (15, 50, 50))
data = [harvester.Gyy(x, y, z, gyy), harvester.Gyz(x, y, z, gyz), harvester.Gzz(x, y, z, gzz)]
seeds = harvester.sow([(xs, 3250, 600, {'density': 1200}) for xs in range(800, 4400, 450)]
                      + [(xs, 2050, 600, {'density': 1200}) for xs in (3300, 3700, 4100)], mesh)
t0 = time.perf_counter()
estimate, predicted, report = harvester.harvest(data, seeds, mesh, compactness=1.0, threshold=0.0001,
                                                report=True)
print("%d accretions in %.2f s; data misfit %.4f" % (report['accretions'], time.perf_counter() - t0,
                                                     report['misfit']))
mesh.addprop('density', estimate['density'])
dens = np.array(list(estimate['density']))
print("%d of %d cells carry density; residual std gzz %.2f Eotvos" % (
    np.count_nonzero(dens), mesh.size, np.std(gzz - predicted[2])))
