"""
Accuracy of the reference and of this repository's evaluators against a 50-digit
(mpmath) evaluation of the same closed-form sums ("report max error vs
mpmath truth for reference and new code on a sample").

    python tools/accuracy_report.py > profiles/r01_accuracy_vs_mpmath.txt

Runs on the CPU: the reference side is the oracle (bit-identical to the compiled
Cython kernel, tests/test_oracle.py), the "new" side is tests/hostsim -- the very
device headers (prism_fast.cuh, prism_mesh.cuh, prism_face.cuh) compiled for the host;
the GPU kernels execute the same code (test_gpu_parity.py holds both to the oracle).
Errors are in units of eps*A (A = sum of |summands|, the parity floor's scale) and
relative to |truth|.
"""
import ctypes
import os
import sys

import mpmath as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_hostsim as th                      # noqa: E402
from fatiando_b200 import mesher, gridder, utils   # noqa: E402
from tools.synthetic import hill   # noqa: E402
from fatiando_b200.flatten import flatten     # noqa: E402
from oracle import prism_oracle as oracle     # noqa: E402

mp.mp.dps = 50
EPS = 2.0 ** -52
H = ctypes.CDLL(os.path.join(ROOT, "tests", "hostsim", "libhostsim.so"))


def sat(y, x):
    if y == 0:
        return mp.mpf(0)
    if x == 0:
        return mp.sign(y) * mp.pi / 2
    return mp.atan(y / x)


def slog(x):
    return mp.mpf(0) if x == 0 else mp.log(x)


def truth(field, xp, yp, zp, bounds, dens):
    """50-digit evaluation of the reference's formula (unshifted; sample points are generic)."""
    out = []
    for x0, y0, z0 in zip(xp, yp, zp):
        tot = mp.mpf(0)
        for q in range(bounds[0].size):
            s = mp.mpf(0)
            for i, xb in enumerate((bounds[1][q], bounds[0][q])):
                for j, yb in enumerate((bounds[3][q], bounds[2][q])):
                    for k, zb in enumerate((bounds[5][q], bounds[4][q])):
                        X, Y, Z = mp.mpf(xb) - mp.mpf(x0), mp.mpf(yb) - mp.mpf(y0), mp.mpf(zb) - mp.mpf(z0)
                        r = mp.sqrt(X * X + Y * Y + Z * Z)
                        if field == 'gz':
                            kk = -(X * slog(Y + r) + Y * slog(X + r) - Z * sat(X * Y, Z * r))
                        elif field == 'gzz':
                            kk = -sat(X * Y, Z * r)
                        elif field == 'gxx':
                            kk = -sat(Z * Y, X * r)
                        elif field == 'gxy':
                            kk = slog(Z + r)
                        s += (-1) ** (i + j + k) * kk
            tot += s * mp.mpf(float(dens[q]))
        out.append(tot)
    return out


def report(title, xp, yp, zp, flat, evaluators, fields=('gz', 'gxx', 'gxy', 'gzz')):
    print("\n## " + title)
    print("%-5s %-22s %14s %14s" % ("field", "evaluator", "max err/(eps*A)", "max rel err"))
    dens = flat.density
    for f in fields:
        t = truth(f, xp, yp, zp, flat.bounds, dens)
        cond = oracle.condition(f, xp, yp, zp, flat.bounds, dens[:, None], scale=1.0)
        tf = np.array([float(v) for v in t])
        for name, fn in evaluators:
            got = fn(f)
            err = np.array([abs(mp.mpf(float(g)) - v) for g, v in zip(got, t)], dtype=float)
            print("%-5s %-22s %14.4f %14.2e" % (f, name, np.max(err / (EPS * cond)),
                                               np.max(err / np.maximum(np.abs(tf), 1e-300))))


def main():
    names = th.NAMES
    rs = np.random.RandomState(3)
    # ---- a regular mesh, near and far points ----
    mesh = mesher.PrismMesh((0, 5000, -300, 4100, 10, 2000), (3, 5, 4))
    mesh.addprop('density', rs.uniform(-1000, 1000, mesh.size))
    flat = flatten(mesh, 'density')
    for label, (xp, yp, zp) in (("PrismMesh 4x5x3, points 150 m above", gridder.scatter((0, 5000, -300, 4100), 12, z=-150., seed=1)),
                                ("PrismMesh 4x5x3, points 60 body sizes away", gridder.scatter((-3e5, 3e5, -3e5, 3e5), 12, z=-4e4, seed=2))):
        def ref(f):
            return oracle.model(f, xp, yp, zp, flat.bounds, flat.density[:, None], scale=1.0)

        def cell(f):
            return th.forward(H, 1 << names.index(f), 0, xp, yp, zp, flat.bounds, flat.density[:, None])[0]

        def node(f):
            return th.mesh_forward(H, 1 << names.index(f), xp, yp, zp, flat.nodes)[0]
        report(label, xp, yp, zp, flat, [("reference (oracle)", ref), ("per cell, fused", cell), ("node sharing", node)])
    # ---- a relief miniature (1 km cells), points 5 km above ----
    area, shape = (0, 12e3, 0, 12e3), (12, 12)
    x, y = gridder.regular(area, shape)
    h = 2500 * hill(x, y, 3e3, 4e3, (5e3, 7e3)) + rs.uniform(0, 200, x.size)
    relief = mesher.PrismRelief(0, gridder.spacing(area, shape)[::-1], (x, y, -h))
    relief.addprop('density', 2670. * np.ones(x.size))
    flat = flatten(relief, 'density')
    xp, yp, zp = gridder.scatter(area, 12, z=-5000., seed=4)

    def ref(f):
        return oracle.model(f, xp, yp, zp, flat.bounds, flat.density[:, None], scale=1.0)

    def cell(f):
        return th.forward(H, 1 << names.index(f), 0, xp, yp, zp, flat.bounds, flat.density[:, None])[0]

    def surf(f):
        return th.surface_forward(H, 1 << names.index(f), xp, yp, zp, flat.surface)[0][0]
    report("PrismRelief 12x12 (1 km cells), points 5 km above", xp, yp, zp, flat,
           [("reference (oracle)", ref), ("per cell, fused", cell), ("surface form", surf)])


if __name__ == "__main__":
    print("# Accuracy against a 50-digit evaluation (tools/accuracy_report.py); eps = 2^-52")
    main()
