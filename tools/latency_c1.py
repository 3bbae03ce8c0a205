"""Latency of the small cookbook case C1 (3 prisms x 100x100 points) through the public API."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fatiando_b200 import prism, mesher, gridder

model = [mesher.Prism(-4000, -3000, -4000, -3000, 0, 2000, {'density': 1000}),
         mesher.Prism(-1000, 1000, -1000, 1000, 0, 2000, {'density': -900}),
         mesher.Prism(2000, 4000, 3000, 4000, 0, 2000, {'density': 1300})]
xp, yp, zp = gridder.regular((-5000, 5000, -5000, 5000), (100, 100), z=-150)
names = list(prism.GRAVITY_FIELDS)
for label, fn in [("prism.gz (one field, one-shot)", lambda: prism.gz(xp, yp, zp, model)),
                  ("prism.fields (10 fields, one-shot)", lambda: prism.fields(xp, yp, zp, model, names))]:
    fn()
    t = []
    for _ in range(50):
        t0 = time.perf_counter(); fn(); t.append(time.perf_counter() - t0)
    print("%-40s median %.3f ms  min %.3f ms" % (label, 1e3 * np.median(t), 1e3 * min(t)))
with prism.Model(model, 'density') as m:
    m.fields(xp, yp, zp, names)
    t = []
    for _ in range(50):
        t0 = time.perf_counter(); m.fields(xp, yp, zp, names); t.append(time.perf_counter() - t0)
    print("%-40s median %.3f ms  min %.3f ms  (kernels %.3f ms)" % ("Model.fields (resident model)", 1e3 * np.median(t), 1e3 * min(t), m.last_kernel_ms()[0]))
