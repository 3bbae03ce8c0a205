"""
Numerics of the fused fast path, studied on the HOST (no GPU needed).

tests/hostsim compiles the device math headers (prism_fast.cuh, prism_exact.cuh)
with g++; this file compares them with the oracle under the project's stated
tolerance.  It validates the ALGEBRA of the collapsed corner sums and the custom
log/atan/sqrt; the same comparisons run on the real kernels in test_gpu_parity.py.
hostsim is test infrastructure: the product never loads it.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden
from helpers import C_FLOOR_FAST, EPS, assert_parity, floor_fast, golden_model

HS_DIR = os.path.join(ROOT, "tests", "hostsim")
dp = ctypes.POINTER(ctypes.c_double)
SETS_G = [0x001, 0x002, 0x004, 0x008, 0x010, 0x020, 0x040, 0x080, 0x100, 0x200, 0x00E, 0x3F0,
          0x3F8, 0x3FE, 0x3FF]
SETS_M = [0x0400, 0x0800, 0x1000, 0x2000, 0x3800, 0x3C00]
NAMES = ("potential", "gx", "gy", "gz", "gxx", "gxy", "gxz", "gyy", "gyz", "gzz",
         "tf", "bx", "by", "bz")


@pytest.fixture(scope="module")
def hostsim():
    subprocess.check_call(["make", "-s", "-C", HS_DIR])
    return ctypes.CDLL(os.path.join(HS_DIR, "libhostsim.so"))


def _p(a):
    return a.ctypes.data_as(dp)


def forward(H, mask, exact, xp, yp, zp, b, props, f=(0, 0, 0)):
    xp, yp, zp = (np.ascontiguousarray(a, dtype=float) for a in (xp, yp, zp))
    b = [np.ascontiguousarray(a, dtype=float) for a in b]
    n, m = xp.size, b[0].size
    out = np.zeros((bin(mask).count('1'), n))
    cols = [np.ascontiguousarray(props[:, i]) for i in range(props.shape[1])]
    barr = (dp * 6)(*[_p(a) for a in b])
    parr = (dp * 3)(*([_p(a) for a in cols] + [None] * (3 - len(cols))))
    fv = np.array(f, dtype=float)
    rc = H.hostsim_forward(ctypes.c_uint32(mask), int(exact), _p(xp), _p(yp), _p(zp),
                           ctypes.c_size_t(n), ctypes.c_size_t(m), barr, parr, _p(fv), _p(out))
    assert rc == 0
    return out


def check_sets(H, oracle, sets, xp, yp, zp, b, props, f=(0, 0, 0), what=''):
    worst = 0.0
    cache = {}
    for mask in sets:
        fast = forward(H, mask, 0, xp, yp, zp, b, props, f)
        exact = forward(H, mask, 1, xp, yp, zp, b, props, f)
        fields = [NAMES[i] for i in range(14) if mask >> i & 1]
        for c, fld in enumerate(fields):
            if fld not in cache:
                pr = props if fld != 'tf' else np.hstack([props, np.tile(f, (props.shape[0], 1))])
                cache[fld] = (oracle.model(fld, xp, yp, zp, b, pr, scale=1.0),
                              oracle.condition(fld, xp, yp, zp, b, pr, scale=1.0))
            ref, cond = cache[fld]
            # host build of the reference-order evaluator: same libm, same order -> same bits
            assert np.array_equal(exact[c], ref), '%s %s exact' % (what, fld)
            worst = max(worst, assert_parity(fast[c], ref, cond, floor_fast(b[0].size),
                                             '%s %s in set 0x%x' % (what, fld, mask)))
    return worst


def test_primitives_against_mpmath(hostsim):
    import mpmath as mp
    mp.mp.dps = 40
    rs = np.random.RandomState(0)
    n = 8000

    def run(name, a, b):
        a, b = np.ascontiguousarray(a, float), np.ascontiguousarray(b, float)
        out = np.empty_like(a)
        getattr(hostsim, name)(_p(a), _p(b), ctypes.c_size_t(a.size), _p(out))
        return out
    a = rs.uniform(-1, 1, n) * 10.0 ** rs.uniform(-20, 20, n)
    b = rs.uniform(0.5, 1, n) * 10.0 ** rs.uniform(-30, 40, n)
    assert np.max(np.abs(run('hostsim_div', a, b) - a / b) / np.spacing(np.abs(a / b))) <= 1.0
    num = rs.uniform(1, 2, n) * 10.0 ** rs.uniform(-10, 30, n)
    den = np.abs(num * np.concatenate([1 + rs.uniform(-1, 1, n // 2) * 10.0 ** rs.uniform(-12, 0, n // 2),
                                       10.0 ** rs.uniform(-16, 16, n // 2)])) + 1e-300
    ex = np.array([float(mp.log(mp.mpf(x) / mp.mpf(y))) for x, y in zip(num, den)])
    assert np.max(np.abs(run('hostsim_log_ratio', num, den) - ex) / np.spacing(np.abs(ex))) <= 2.0
    re = np.abs(rs.normal(size=n)) * 10.0 ** rs.uniform(-5, 20, n)
    im = rs.normal(size=n) * 10.0 ** rs.uniform(-5, 20, n)
    re[:50] = 0.0
    ex = np.array([float(mp.atan2(mp.mpf(y), mp.mpf(x))) for x, y in zip(re, im)])
    assert np.max(np.abs(run('hostsim_atan_rhp', im, re) - ex) / np.spacing(np.abs(ex))) <= 2.0


def test_near_field_random_model(hostsim, oracle):
    rs = np.random.RandomState(0)
    n, m = 600, 25
    xp, yp, zp = rs.uniform(-5e3, 5e3, n), rs.uniform(-5e3, 5e3, n), rs.uniform(-500, 1500, n)
    x1 = rs.uniform(-4e3, 3e3, m); x2 = x1 + rs.uniform(10, 1500, m)
    y1 = rs.uniform(-4e3, 3e3, m); y2 = y1 + rs.uniform(10, 1500, m)
    z1 = rs.uniform(0, 1e3, m); z2 = z1 + rs.uniform(10, 1500, m)
    b = (x1, x2, y1, y2, z1, z2)
    w = check_sets(hostsim, oracle, SETS_G, xp, yp, zp, b, rs.uniform(-1e3, 1e3, (m, 1)), what='near')
    assert w <= 1.0
    check_sets(hostsim, oracle, SETS_M, xp, yp, zp, b, rs.uniform(-2, 2, (m, 3)), (0.3, 0.5, 0.81), 'near')


def test_points_on_and_almost_on_faces_edges_vertices(hostsim, oracle):
    """Exact zeros (aligned grids) and near-zeros where sqrt() absorbs the offset."""
    rs = np.random.RandomState(4)
    xe = np.linspace(0, 600, 7)
    I, J, K = np.meshgrid(range(6), range(6), range(3), indexing='ij')
    b = [xe[I.ravel()], xe[I.ravel() + 1], xe[J.ravel()], xe[J.ravel() + 1],
         100.0 * K.ravel(), 100.0 * K.ravel() + 100.0]
    m = b[0].size
    g = np.linspace(-100, 700, 9)
    PX, PY, PZ = np.meshgrid(g, g, np.array([-100., 0., 100., 150., 300., 400.]), indexing='ij')
    xp, yp, zp = PX.ravel().copy(), PY.ravel().copy(), PZ.ravel().copy()
    # nudge a third of the points by relative 1e-16 .. 1e-7: almost aligned
    sel = rs.rand(xp.size) < 0.33
    xp[sel] += xp[sel] * 10.0 ** rs.uniform(-16, -7, sel.sum())
    sel = rs.rand(xp.size) < 0.33
    yp[sel] -= (yp[sel] + 1.0) * 10.0 ** rs.uniform(-16, -7, sel.sum())
    sel = rs.rand(xp.size) < 0.2
    zp[sel] += (zp[sel] + 1.0) * 10.0 ** rs.uniform(-16, -7, sel.sum())
    check_sets(hostsim, oracle, [0x3FF, 0x008, 0x3F0, 0x020, 0x100, 0x040], xp, yp, zp, b,
               rs.uniform(-1e3, 1e3, (m, 1)), what='aligned')
    check_sets(hostsim, oracle, [0x3C00, 0x0400, 0x0800], xp, yp, zp, b, rs.uniform(-2, 2, (m, 3)),
               (0.3, 0.5, 0.81), 'aligned')
    d = load_golden("edge_cases")
    bounds, props = golden_model(d, 'gz')
    check_sets(hostsim, oracle, SETS_G, d['xp'], d['yp'], d['zp'], bounds, props, what='edge')


def test_far_field_relief(hostsim, oracle):
    rs = np.random.RandomState(3)
    xs = np.linspace(0, 1e6, 40)
    X, Y = np.meshgrid(xs, xs, indexing='ij')
    xc, yc = X.ravel(), Y.ravel()
    dx = xs[1] - xs[0]
    h = 3000 * np.exp(-((xc - 4e5) ** 2 + (yc - 6e5) ** 2) / 2e5 ** 2) + rs.uniform(0, 200, xc.size) - 500
    zc = -h
    b = (xc - 0.5 * dx, xc + 0.5 * dx, yc - 0.5 * dx, yc + 0.5 * dx,
         np.where(zc <= 0, zc, 0.0), np.where(zc <= 0, 0.0, zc))
    dens = np.where(zc > 0, -2670.0, 2670.0)[:, None]
    g = np.linspace(0, 1e6, 12)
    PX, PY = np.meshgrid(g, g, indexing='ij')
    xp, yp = PX.ravel().copy(), PY.ravel().copy()
    zp = np.full(xp.size, -5000.0)
    w = check_sets(hostsim, oracle, [0x008, 0x3F8], xp, yp, zp, b, dens, what='relief')
    assert w <= 1.0
    # how tight it really is: the far-field discrepancy is a small fraction of eps*A
    fast = forward(hostsim, 0x008, 0, xp, yp, zp, b, dens)[0]
    ref = oracle.model('gz', xp, yp, zp, b, dens, scale=1.0)
    cond = oracle.condition('gz', xp, yp, zp, b, dens, scale=1.0)
    assert np.max(np.abs(fast - ref) / (EPS * cond)) < 0.25


def mesh_forward(H, mask, xp, yp, zp, nodes, f=(0, 0, 0)):
    xp, yp, zp = (np.ascontiguousarray(a, dtype=float) for a in (xp, yp, zp))
    out = np.zeros((bin(mask).count('1'), xp.size))
    w = [np.ascontiguousarray(a) for a in nodes.weights]
    warr = (dp * 3)(*([_p(a) for a in w] + [None] * (3 - len(w))))
    sh = 0.00001 * nodes.cell_size
    fv = np.array(f, dtype=float)
    rc = H.hostsim_mesh_forward(ctypes.c_uint32(mask), _p(xp), _p(yp), _p(zp), ctypes.c_size_t(xp.size),
                                nodes.xn.size, nodes.yn.size, nodes.zn.size, _p(nodes.xn),
                                _p(nodes.yn), _p(nodes.zn), warr, _p(sh), _p(fv), _p(out))
    assert rc == 0
    return out


def test_node_sharing_matches_the_per_cell_reference(hostsim, oracle):
    """Regular meshes evaluated once per NODE (csrc/prism_mesh.cuh) against the oracle's
    per-cell loop: plain and awkward bounds, masked cells, grids on the mesh planes,
    points inside the mesh, far points; gravity sets and magnetic sets."""
    from fatiando_b200 import mesher, gridder, utils
    from fatiando_b200.flatten import flatten, mesh_nodes
    rs = np.random.RandomState(12)
    cases = [((4, 6, 7), (0, 5000, 0, 5000, 0, 2000), -150.0, None),
             ((3, 7, 5), (0.1, 5000.3, -0.7, 4999.9, 0.05, 2000.95), -150.0, [0, 9, 40, 104]),
             ((4, 6, 7), (0, 5000, 0, 5000, 0, 2000), -60000.0, [3, 50]),
             ((2, 5, 5), (0, 500, 0, 500, 0, 200), 100.0, None),          # grid through the mesh
             # seen from 60 body sizes away on every side: on the +x / +y side the reference
             # folds atan2 by pi (oracle abs_atan_terms) and is itself the inaccurate one
             ((3, 4, 5), (0, 5000, -300, 4100, 10, 2000), 'far', [7])]
    for shape, bounds, z0, mask in cases:
        mesh = mesher.PrismMesh(bounds, shape)
        mesh.addprop('density', rs.uniform(-1000, 1000, mesh.size))
        mesh.addprop('magnetization', rs.uniform(-2, 2, (mesh.size, 3)))
        if mask:
            mesh.mask = mask
        nx, ny = shape[2], shape[1]
        # every second grid line lies in a mesh plane
        if z0 == 'far':
            xp, yp, zp = gridder.regular((-3e5, 3e5, -3e5, 3e5), (9, 7), z=-4e4)
        else:
            xp, yp, zp = gridder.regular((bounds[0], bounds[1], bounds[2], bounds[3]),
                                         (2 * nx + 1, 2 * ny + 1), z=z0)
        for prop, sets, f in (('density', [0x3FF, 0x008, 0x3F0, 0x020], (0, 0, 0)),
                              ('magnetization', [0x3C00, 0x0400], utils.dircos(30, -15))):
            flat = flatten(mesh, prop)
            nodes = mesh_nodes(mesh, prop)
            props = flat.density[:, None] if prop == 'density' else flat.magnetization
            for m in sets:
                got = mesh_forward(hostsim, m, xp, yp, zp, nodes, f)
                fields = [NAMES[i] for i in range(14) if m >> i & 1]
                for c, fld in enumerate(fields):
                    pr = props if fld != 'tf' else np.hstack([props, np.tile(f, (props.shape[0], 1))])
                    ref = oracle.model(fld, xp, yp, zp, flat.bounds, pr, scale=1.0)
                    cond = oracle.condition(fld, xp, yp, zp, flat.bounds, pr, scale=1.0)
                    assert_parity(got[c], ref, cond, C_FLOOR_FAST,
                                  'nodes %s %s z=%s set 0x%x' % (shape, fld, z0, m))
    # uniform density: interior node weights vanish, only the hull is evaluated
    mesh = mesher.PrismMesh((0, 1000, 0, 1000, 0, 500), (5, 6, 7))
    mesh.addprop('density', 300.0 * np.ones(mesh.size))
    nodes = mesh_nodes(mesh, 'density')
    assert np.count_nonzero(nodes.weights[0]) == 8


def surface_forward(H, mask, xp, yp, zp, surf):
    xp, yp, zp = (np.ascontiguousarray(a, dtype=float) for a in (xp, yp, zp))
    out = np.zeros((bin(mask).count('1'), xp.size))
    farr = (dp * 7)(*[_p(a) for a in surf.faces])
    narr = (dp * 4)(*[_p(a) for a in surf.nodes])
    flag = H.hostsim_surface_forward(ctypes.c_uint32(mask), _p(xp), _p(yp), _p(zp),
                                     ctypes.c_size_t(xp.size), ctypes.c_size_t(surf.faces[0].size), farr,
                                     ctypes.c_size_t(surf.nodes[0].size), narr, _p(surf.cell_size), _p(out))
    assert flag >= 0
    return out, flag


def test_relief_surface_form_matches_the_per_cell_reference(hostsim, oracle):
    """A PrismRelief as top faces + merged reference-plane nodes (csrc/prism_face.cuh) against
    the oracle's per-cell loop: relief crossing the reference level, uniform and variable
    density, points above, far away, between the cells' heights and on grid lines."""
    from fatiando_b200 import mesher, gridder
    from fatiando_b200.flatten import flatten, relief_surface
    rs = np.random.RandomState(8)
    area, shape = (0, 8000, -1000, 5000), (9, 11)
    x, y = gridder.regular(area, shape)
    h = 600 * np.sin(x / 900.) * np.cos(y / 700.) + rs.uniform(-80, 80, x.size)     # crosses ref = 0
    sets = [0x3FF, 0x008, 0x3F8, 0x3F0, 0x00E, 0x040, 0x100]
    for variable in (False, True):
        # dims = (dy, dx) (mesh.py:429)
        relief = mesher.PrismRelief(0, gridder.spacing(area, shape)[::-1], (x, y, -h))
        relief.addprop('density', rs.uniform(2000, 3000, x.size) if variable else 2670. * np.ones(x.size))
        surf = relief_surface(relief)
        assert surf is not None
        if not variable:
            # uniform density: interior reference-plane nodes cancel, the rim stays
            assert surf.nodes[0].size <= 2 * (shape[0] + 1) + 2 * (shape[1] + 1)
        flat = flatten(relief, 'density')
        assert flat.surface is not None
        grids = [gridder.regular(area, (7, 6), z=-1500.),
                 gridder.regular((-4e5, 4e5, -4e5, 4e5), (5, 6), z=-3e4),
                 gridder.regular(area, (9, 11), z=-700.),                 # on the cell centres, above
                 gridder.scatter(area, 40, z=100., seed=3),                # below some of the tops
                 (surf.faces[0][:30] + 0.0, surf.faces[2][:30] + 0.0, -900. * np.ones(30))]   # over corners
        for xp, yp, zp in grids:
            for m in sets:
                got, flag = surface_forward(hostsim, m, xp, yp, zp, surf)
                assert flag == 0
                fields = [NAMES[i] for i in range(14) if m >> i & 1]
                for c, fld in enumerate(fields):
                    ref = oracle.model(fld, xp, yp, zp, flat.bounds, flat.density[:, None], scale=1.0)
                    cond = oracle.condition(fld, xp, yp, zp, flat.bounds, flat.density[:, None], scale=1.0)
                    assert_parity(got[c], ref, cond, C_FLOOR_FAST, 'relief surface %s set 0x%x' % (fld, m))
    # a point IN the reference plane, on a grid line, beyond a node: gxz / gyz raise the flag
    xp = np.array([surf.nodes[0][0]]); yp = np.array([surf.nodes[1][0] + 7000.]); zp = np.zeros(1)
    _, flag = surface_forward(hostsim, 0x040, xp, yp, zp, surf)
    assert flag == 1
    _, flag = surface_forward(hostsim, 0x008, xp, yp, zp, surf)
    assert flag == 0
    # not a tight regular grid: no surface form
    relief = mesher.PrismRelief(0, (500., 500.), (x, y, -h))
    relief.addprop('density', 2670. * np.ones(x.size))
    assert relief_surface(relief) is None
    xs, ys = gridder.scatter(area, 50, seed=0)
    relief = mesher.PrismRelief(0, (100., 100.), (xs, ys, -np.ones(50)))
    relief.addprop('density', 2670. * np.ones(50))
    assert relief_surface(relief) is None


def test_accuracy_against_a_50_digit_evaluation(hostsim, oracle):
    """the new evaluators are no less accurate than the reference.  Truth = mpmath
    evaluation of the same sums (tools/accuracy_report.py); every evaluator, and the reference
    itself, stays within a quarter of eps*A of it on a mesh seen from near and from afar."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import accuracy_report as ar
    from fatiando_b200 import mesher, gridder
    from fatiando_b200.flatten import flatten
    rs = np.random.RandomState(3)
    mesh = mesher.PrismMesh((0, 5000, -300, 4100, 10, 2000), (2, 3, 3))
    mesh.addprop('density', rs.uniform(-1000, 1000, mesh.size))
    flat = flatten(mesh, 'density')
    for xp, yp, zp in (gridder.scatter((0, 5000, -300, 4100), 5, z=-150., seed=1),
                       gridder.scatter((-3e5, 3e5, -3e5, 3e5), 5, z=-4e4, seed=2)):
        for f in ('gz', 'gxx', 'gxy'):
            t = ar.truth(f, xp, yp, zp, flat.bounds, flat.density)
            cond = oracle.condition(f, xp, yp, zp, flat.bounds, flat.density[:, None], scale=1.0)
            got = {'reference': oracle.model(f, xp, yp, zp, flat.bounds, flat.density[:, None], scale=1.0),
                   'per cell': forward(hostsim, 1 << NAMES.index(f), 0, xp, yp, zp, flat.bounds,
                                       flat.density[:, None])[0],
                   'nodes': mesh_forward(hostsim, 1 << NAMES.index(f), xp, yp, zp, flat.nodes)[0]}
            for name, vals in got.items():
                err = np.array([abs(ar.mp.mpf(float(v)) - tv) for v, tv in zip(vals, t)], dtype=float)
                assert np.max(err / (EPS * cond)) < 0.25, (f, name)
