import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from fatiando_b200 import prism
w = bench.make_workload('c2', 1)
xp, yp, zp = w['xp'], w['yp'], w['zp']
for _ in range(3):
    prism.fields(xp, yp, zp, w['model'], w['names'])
t = []
for _ in range(10):
    t0 = time.perf_counter(); prism.fields(xp, yp, zp, w['model'], w['names']); t.append(time.perf_counter() - t0)
print('e2e ms', 1e3 * np.median(t))
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    prism.fields(xp, yp, zp, w['model'], w['names'])
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
