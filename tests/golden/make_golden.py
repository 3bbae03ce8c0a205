"""
Generate the golden fixtures of tests/golden/ from the REAL reference.

Run in the build container only (needs /root/reference and the compiled
oracle/_ref extension):  python tests/golden/make_golden.py

It imports the reference's own modules -- fatiando/gravmag/prism.py (driving
the unmodified Cython _prism built by oracle/build_ref.py), mesher/geometry.py,
mesher/mesh.py, utils.py, gridder/point_generation.py -- straight from
/root/reference, without running the package __init__ files (they pull in
matplotlib/py2-only code that is unrelated to this path).  Shims, all of them
Python-2-isms of the reference, none touching arithmetic:
  * numpy.float = float                      (_prism.pyx:13, prism.py:130, ...)
  * numpy.meshgrid returns a list            (point_generation.py:93-95 appends)
  * stub `future.builtins`; its `object` maps __next__ to .next (mesh.py:460,642)
Nothing from this repo's product package is imported here: the fixtures are
pure reference outputs.  Inputs are stored next to outputs so that the GPU box
(which has no /root/reference) can replay them.
"""
import glob
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def import_reference():
    np.float = float
    _meshgrid = np.meshgrid
    np.meshgrid = lambda *a, **k: list(_meshgrid(*a, **k))

    class _Object(object):
        def __next__(self):
            return self.next()

    fut = types.ModuleType("future")
    fb = types.ModuleType("future.builtins")
    fb.range, fb.super, fb.zip, fb.map, fb.isinstance = range, super, zip, map, isinstance
    fb.object = _Object
    fut.builtins = fb
    sys.modules["future"] = fut
    sys.modules["future.builtins"] = fb

    def pkg(name, paths):
        m = types.ModuleType(name)
        m.__path__ = paths
        sys.modules[name] = m
        return m

    so_dir = os.path.join(ROOT, "oracle", "_ref")
    if not glob.glob(os.path.join(so_dir, "_prism*.so")):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import build_ref
        build_ref.build()
    fat = pkg("fatiando", [os.path.join(REF, "fatiando")])
    pkg("fatiando.gravmag", [os.path.join(REF, "fatiando", "gravmag"), so_dir])
    pkg("fatiando.mesher", [os.path.join(REF, "fatiando", "mesher")])
    grd = pkg("fatiando.gridder", [os.path.join(REF, "fatiando", "gridder")])
    import importlib
    pg = importlib.import_module("fatiando.gridder.point_generation")
    gu = importlib.import_module("fatiando.gridder.utils")
    grd.regular, grd.scatter, grd.spacing = pg.regular, pg.scatter, gu.spacing
    fat.gridder = grd
    fat.constants = importlib.import_module("fatiando.constants")
    utils = importlib.import_module("fatiando.utils")
    geometry = importlib.import_module("fatiando.mesher.geometry")
    mesh = importlib.import_module("fatiando.mesher.mesh")
    prism = importlib.import_module("fatiando.gravmag.prism")
    assert prism._prism is not None, "reference Cython kernel not loaded"
    return prism, geometry, mesh, utils, grd


GRAV = ['potential', 'gx', 'gy', 'gz', 'gxx', 'gxy', 'gxz', 'gyy', 'gyz', 'gzz']
MAG = ['tf', 'bx', 'by', 'bz']
KERN = ['xx', 'xy', 'xz', 'yy', 'yz', 'zz']


def prisms_to_arrays(prisms):
    """Reference iteration result, as plain arrays (None -> nan row, flag)."""
    rows, dens, mag, isnone = [], [], [], []
    for p in prisms:
        if p is None:
            rows.append([np.nan] * 6)
            dens.append(np.nan)
            mag.append([np.nan] * 3)
            isnone.append(True)
            continue
        isnone.append(False)
        rows.append(p.get_bounds())
        dens.append(p.props.get('density', np.nan))
        m = p.props.get('magnetization', [np.nan] * 3)
        mag.append(list(np.atleast_1d(m)) if np.ndim(m) else [m, np.nan, np.nan])
    return dict(bounds=np.array(rows, dtype=float).reshape(-1, 6),
                density=np.array(dens, dtype=float),
                magnetization=np.array(mag, dtype=float).reshape(-1, 3),
                isnone=np.array(isnone, dtype=bool))


def main():
    prism, geometry, mesh, utils, gridder = import_reference()
    Prism, PrismMesh, PrismRelief = geometry.Prism, mesh.PrismMesh, mesh.PrismRelief
    out = {}

    # ---- A. test_prism.py:198-234 model on a 41x41 version of its grid -----
    inc, dec = -30, 50
    model = [Prism(100, 300, -100, 100, 0, 400,
                   {'density': -1000, 'magnetization': utils.ang2vec(-2, inc, dec)}),
             Prism(-300, -100, -100, 100, 0, 200,
                   {'density': 2000, 'magnetization': utils.ang2vec(5, 25, -10)})]
    tmp = np.linspace(-500, 500, 41)
    xp, yp = [i.ravel() for i in np.meshgrid(tmp, tmp)]
    zp = -1 * np.ones_like(xp)
    d = dict(xp=xp, yp=yp, zp=zp, inc=inc, dec=dec, **prisms_to_arrays(model))
    for f in GRAV + ['bx', 'by', 'bz']:
        d[f] = getattr(prism, f)(xp, yp, zp, model)
    d['tf'] = prism.tf(xp, yp, zp, model, inc, dec)
    for c in KERN:
        for q, p in enumerate(model):
            d['kernel%s_%d' % (c, q)] = getattr(prism, 'kernel' + c)(xp, yp, zp, p)
    out['two_prisms'] = d

    # ---- B. config C1: cookbook/gravmag_grav_prism.py:9-13 verbatim --------
    model = [Prism(-4000, -3000, -4000, -3000, 0, 2000, {'density': 1000}),
             Prism(-1000, 1000, -1000, 1000, 0, 2000, {'density': -900}),
             Prism(2000, 4000, 3000, 4000, 0, 2000, {'density': 1300})]
    xp, yp, zp = gridder.regular((-5000, 5000, -5000, 5000), (100, 100), z=-150)
    d = dict(xp=xp, yp=yp, zp=zp, **prisms_to_arrays(model))
    for f in GRAV:
        d[f] = getattr(prism, f)(xp, yp, zp, model)
    out['c1_cookbook'] = d

    # ---- C. edge cases: SURVEY App. F prism, points on faces/edges/vertices,
    # inside, aligned below/east/north (singularity shift active) -----------
    inc, dec = 30, -10
    model = [Prism(-100, 100, -200, 200, 50, 300,
                   {'density': 1000, 'magnetization': utils.ang2vec(2, inc, dec)})]
    pts = np.array([
        (100, 200, 0), (100, 200, 400), (0, 0, 50), (0, 0, 175), (100, 200, 175),
        (100, 200, 50), (-100, -200, 300), (-100, 200, 500), (100, -200, 1000),
        (100, 500, 50), (100, 500, 300), (-100, 700, 50),     # gxz shift (dy<0)
        (400, 200, 50), (400, -200, 300), (900, 200, 300),    # gyz shift (dx<0)
        (0, 0, 0), (0, 0, 300), (0, 200, 100), (100, 0, 100), (50, 60, 70),
        (-100, 0, 50), (100, 200, 300), (1e6, 2e6, -3e5), (-3e4, 100, 200),
        (100, 200, -1e5), (0, 0, -1e-3), (100.0000001, 200, 175)], dtype=float)
    xp, yp, zp = pts[:, 0].copy(), pts[:, 1].copy(), pts[:, 2].copy()
    d = dict(xp=xp, yp=yp, zp=zp, inc=inc, dec=dec, **prisms_to_arrays(model))
    for f in GRAV + ['bx', 'by', 'bz']:
        d[f] = getattr(prism, f)(xp, yp, zp, model)
    d['tf'] = prism.tf(xp, yp, zp, model, inc, dec)
    for c in KERN:
        d['kernel%s_0' % c] = getattr(prism, 'kernel' + c)(xp, yp, zp, model[0])
    # scalar magnetization (induced along F) -- prism.py:650-653
    smodel = [Prism(-100, 100, -200, 200, 50, 300, {'magnetization': 2.0})]
    d['tf_scalar_mag'] = prism.tf(xp, yp, zp, smodel, inc, dec)
    # inverted and zero-thickness prisms
    d['gz_inverted'] = prism.gz(xp, yp, zp, [Prism(100, -100, -200, 200, 50, 300, {'density': 1000})])
    d['gz_flat'] = prism.gz(xp, yp, zp, [Prism(-100, 100, -200, 200, 50, 50, {'density': 1000})])
    out['edge_cases'] = d

    # ---- D. test_prism.py:237-250 cube, six faces, 21x21 (hits x,y=+-300) ---
    model = [Prism(-300, 300, -300, 300, -300, 300, {'density': 1000})]
    shape, area, distance = (21, 21), [-600, 600, -600, 600], 310
    grids = [gridder.regular(area, shape, z=-distance),
             gridder.regular(area, shape, z=distance),
             gridder.regular(area, shape, z=distance)[::-1],
             gridder.regular(area, shape, z=-distance)[::-1],
             np.array(gridder.regular(area, shape, z=distance))[[0, 2, 1]],
             np.array(gridder.regular(area, shape, z=-distance))[[0, 2, 1]]]
    d = dict(**prisms_to_arrays(model))
    for g, (x, y, z) in enumerate(grids):
        x, y, z = np.array(x), np.array(y), np.array(z)
        d['xp_%d' % g], d['yp_%d' % g], d['zp_%d' % g] = x, y, z
        for f in GRAV:
            d['%s_%d' % (f, g)] = getattr(prism, f)(x, y, z, model)
    out['cube_faces'] = d

    # ---- E. property overrides and skipping (test_prism.py:121-195) --------
    inc, dec = 50, -30
    model = [None,
             Prism(-6000, -2000, 2000, 4000, 0, 3000,
                   {'density': 1000, 'magnetization': utils.ang2vec(10, inc, dec)}),
             Prism(2000, 6000, 2000, 4000, 0, 1000,
                   {'magnetization': utils.ang2vec(15, inc, dec)}),
             None,
             Prism(-6000, -2000, -4000, -2000, 500, 2000, {'density': -1000})]
    xp, yp, zp = gridder.regular([-10000, 10000, -5000, 5000], (21, 11), z=-1)
    d = dict(xp=xp, yp=yp, zp=zp, inc=inc, dec=dec, **prisms_to_arrays(model))
    mag = utils.ang2vec(-5, -30, 15)
    d['pmag'] = np.array(mag)
    for f in GRAV:
        d[f] = getattr(prism, f)(xp, yp, zp, model)
        d[f + '_dens'] = getattr(prism, f)(xp, yp, zp, model, dens=-500)
    for f in ['bx', 'by', 'bz']:
        d[f] = getattr(prism, f)(xp, yp, zp, model)
        d[f + '_pmag'] = getattr(prism, f)(xp, yp, zp, model, pmag=mag)
    d['tf'] = prism.tf(xp, yp, zp, model, inc, dec)
    d['tf_pmag'] = prism.tf(xp, yp, zp, model, inc, dec, pmag=mag)
    d['tf_pmag_scalar'] = prism.tf(xp, yp, zp, model, inc, dec, pmag=3.5)
    out['skip_override'] = d

    # ---- F. PrismMesh: ordering, mask, props; mini C2 (tensor) -------------
    rs = np.random.RandomState(0)
    m = PrismMesh((0, 5000, -300, 4100, 10, 2000), (3, 5, 7))
    m.addprop('density', rs.uniform(-1000, 1000, m.size))
    m.mask = [0, 4, 17, 50, 104]
    xp, yp, zp = gridder.regular((-500, 5500, -500, 4500), (13, 11), z=-150)
    d = dict(xp=xp, yp=yp, zp=zp, mesh_bounds=np.array(m.bounds, float),
             mesh_shape=np.array(m.shape), mesh_mask=np.array(m.mask),
             mesh_density=np.array(m.props['density']),
             mesh_dims=np.array(m.dims, float), **prisms_to_arrays(m))
    for f in GRAV:
        d[f] = getattr(prism, f)(xp, yp, zp, m)
    # awkward bounds: dims not representable, exercises x1+dx vs b0+dx*(i+1)
    m2 = PrismMesh((0.1, 1.3, -0.7, 2.9, 0.05, 0.95), (3, 7, 11))
    d2 = prisms_to_arrays(m2)
    d['awkward_bounds'] = d2['bounds']
    d['awkward_mesh_bounds'] = np.array(m2.bounds, float)
    d['awkward_mesh_shape'] = np.array(m2.shape)
    out['prism_mesh'] = d

    # ---- G. PrismMesh with vector magnetization; mini C3 -------------------
    rs = np.random.RandomState(1)
    m = PrismMesh((0, 10000, 0, 10000, 0, 2000), (2, 6, 5))
    inc, dec = 30, -15
    mi = rs.uniform(0.5, 2, m.size)
    ri = rs.uniform(0, 1, m.size)
    ii = rs.uniform(-90, 90, m.size)
    di = rs.uniform(-180, 180, m.size)
    magv = [np.array(utils.ang2vec(mi[q], inc, dec)) + np.array(utils.ang2vec(ri[q], ii[q], di[q]))
            for q in range(m.size)]
    m.addprop('magnetization', magv)
    xp, yp, zp = gridder.regular((0, 10000, 0, 10000), (12, 14), z=-300)
    d = dict(xp=xp, yp=yp, zp=zp, inc=inc, dec=dec, mesh_bounds=np.array(m.bounds, float),
             mesh_shape=np.array(m.shape), mesh_magnetization=np.array(magv),
             **prisms_to_arrays(m))
    for f in ['bx', 'by', 'bz']:
        d[f] = getattr(prism, f)(xp, yp, zp, m)
    d['tf'] = prism.tf(xp, yp, zp, m, inc, dec)
    out['mesh_magnetic'] = d

    # ---- H. PrismRelief: sign rule of addprop; mini C5 (far-field gz) ------
    rs = np.random.RandomState(2)
    area, shape = (0, 64000, 0, 48000), (16, 12)
    x, y = gridder.regular(area, shape)
    h = 3000 * utils.gaussian2d(x, y, 12000, 9000, 30000, 20000, angle=20) \
        - 1500 * utils.gaussian2d(x, y, 8000, 8000, 50000, 35000) \
        + rs.uniform(0, 200, x.size) - 300
    nodes = (x, y, -h)
    rel = PrismRelief(0, gridder.spacing(area, shape), nodes)
    rel.addprop('density', 2670 * np.ones(rel.size))
    xp, yp, zp = gridder.regular(area, (9, 7), z=-5000)
    d = dict(xp=xp, yp=yp, zp=zp, relief_ref=0.0,
             relief_dims=np.array(gridder.spacing(area, shape), float),
             relief_x=x, relief_y=y, relief_z=-h, relief_density_in=2670 * np.ones(rel.size),
             **prisms_to_arrays(rel))
    for f in ['gz', 'gzz', 'gxy', 'potential']:
        d[f] = getattr(prism, f)(xp, yp, zp, rel)
    out['prism_relief'] = d

    # ---- I. Jacobian, the way the reference builds it (mini C4) ------------
    m = PrismMesh((0, 10000, 0, 10000, 0, 4000), (2, 3, 4))
    xp, yp, zp = gridder.regular((0, 10000, 0, 10000), (6, 5), z=-100)
    inc, dec = 30, -15
    d = dict(xp=xp, yp=yp, zp=zp, inc=inc, dec=dec, mesh_bounds=np.array(m.bounds, float),
             mesh_shape=np.array(m.shape), **prisms_to_arrays(m))
    for f in ['gz', 'gzz', 'gxy', 'potential']:
        jac = np.empty((xp.size, m.size))
        for q, c in enumerate(m):                      # eqlayer.py:108-111 pattern
            jac[:, q] = getattr(prism, f)(xp, yp, zp, [c], dens=1.)
        d['jac_' + f] = jac
    jac = np.empty((xp.size, m.size))
    pm = utils.ang2vec(1, inc, dec)                   # harvester.py:958-962 pattern
    for q, c in enumerate(m):
        jac[:, q] = prism.tf(xp, yp, zp, [c], inc, dec, pmag=pm)
    d['jac_tf'] = jac
    d['jac_tf_pmag'] = np.array(pm)
    out['jacobian'] = d

    # ---- J. support functions ----------------------------------------------
    d = {}
    angs = [(30, -15), (-30, 50), (90, 0), (0, 180), (12.5, 371.25), (-90, 45)]
    d['angles'] = np.array(angs, float)
    d['dircos'] = np.array([utils.dircos(i, dd) for i, dd in angs])
    d['ang2vec_2p5'] = np.array([utils.ang2vec(2.5, i, dd) for i, dd in angs])
    d['ang2vec_array'] = np.array(utils.ang2vec(np.arange(4.), 45, 45))
    x, y, z = gridder.regular((0, 10, -3, 5), (5, 3), z=-10)
    d['regular_x'], d['regular_y'], d['regular_z'] = x, y, z
    d['spacing'] = np.array(gridder.spacing((0, 10, 0, 20), (5, 21)), float)
    C = sys.modules['fatiando.constants']
    d['constants'] = np.array([C.G, C.SI2MGAL, C.SI2EOTVOS, C.CM, C.T2NT,
                               C.G * C.SI2MGAL, C.G * C.SI2EOTVOS, C.CM * C.T2NT])
    out['support'] = d

    for name, d in out.items():
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **{k: np.asarray(v) for k, v in d.items()})
        print('%-16s %4d arrays %8.1f KB' % (name, len(d), os.path.getsize(path) / 1e3))


if __name__ == '__main__':
    main()
