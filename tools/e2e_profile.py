"""
Where the host-buffer (e2e) step of C2 spends its time: cProfile over prism.fields calls,
ctypes entry points shown as their own lines.  python tools/e2e_profile.py [workload] [reps]
"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fatiando_b200 import prism, _lib  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    w = bench.make_workload(name, 1)
    xp, yp, zp = w['xp'], w['yp'], w['zp']

    def step():
        return prism.fields(xp, yp, zp, w['model'], w['names'], inc=w.get('inc'), dec=w.get('dec'))
    for _ in range(3):
        step()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    wall = (time.perf_counter() - t0) / reps
    # wrap the C entry points so that they show up by name
    lib = _lib.lib()
    names = ["fb200_model_create", "fb200_model_attach_mesh", "fb200_model_attach_surface",
             "fb200_model_set_hazard_planes", "fb200_forward", "fb200_model_destroy"]
    timers = {}

    def wrap(nm):
        fn = getattr(lib, nm)

        def inner(*a):
            t = time.perf_counter()
            r = fn(*a)
            timers[nm] = timers.get(nm, 0.0) + time.perf_counter() - t
            return r
        inner.__name__ = nm
        setattr(lib, nm, inner)
    for nm in names:
        wrap(nm)
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    wall2 = (time.perf_counter() - t0) / reps
    print("%s: %.3f ms per prism.fields call (%.3f with the timers)" % (name, wall * 1e3, wall2 * 1e3))
    for nm in names:
        if nm in timers:
            print("  %-32s %.3f ms" % (nm, timers[nm] / reps * 1e3))
    print("  %-32s %.3f ms" % ("python outside the library", (wall2 - sum(timers.values()) / reps) * 1e3))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(reps):
        step()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)


if __name__ == "__main__":
    main()
