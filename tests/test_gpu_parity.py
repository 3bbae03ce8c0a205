"""
Parity of the CUDA path with the oracle and the golden vectors (needs a GPU).

Everything goes through the C ABI (ctypes -> libfatiando_b200.so).  Tolerance
(helpers.py): |new - ref| <= 1e-9*|ref| + c*eps*A with A the sum of
absolute terms from the instrumented oracle; c = helpers.C_FLOOR_FAST for the fused
fast path and the merged forms, c = helpers.C_FLOOR_EXACT for the reference-order path.
"""
import numpy as np
import pytest

from conftest import load_golden
from helpers import (GRAV, MAG, KERN, C_FLOOR_FAST, C_FLOOR_EXACT, EPS, assert_parity, floor_fast, golden_model,
                     prisms_from_golden)

pytestmark = pytest.mark.gpu


def _fvec(d):
    from fatiando_b200 import utils
    return utils.dircos(float(d['inc']), float(d['dec']))


def _cond(oracle, f, d, fvec=None, scale=None):
    bounds, props = golden_model(d, f, fvec)
    return oracle.condition(f, d['xp'], d['yp'], d['zp'], bounds, props, scale=scale)


@pytest.mark.parametrize("case", ["two_prisms", "c1_cookbook", "edge_cases", "skip_override"])
def test_public_api_against_golden(gpu_lib, oracle, case):
    """The drop-in functions on the reference's own test models."""
    from fatiando_b200 import prism, mesher
    d = load_golden(case)
    model = prisms_from_golden(d, mesher.Prism)
    fvec = _fvec(d) if 'inc' in d else None
    xp, yp, zp = d['xp'], d['yp'], d['zp']
    worst = 0.0
    few = floor_fast(int((~d['isnone']).sum()))      # models of 1-3 prisms: see helpers.floor_fast
    for f in GRAV + MAG:
        if f not in d:
            continue
        if f == 'tf':
            got = prism.tf(xp, yp, zp, model, float(d['inc']), float(d['dec']))
        else:
            got = getattr(prism, f)(xp, yp, zp, model)
        worst = max(worst, assert_parity(got, d[f], _cond(oracle, f, d, fvec), few,
                                         '%s/%s' % (case, f)))
    assert worst > 0 or case == 'edge_cases'
    if case == "two_prisms":
        for c in KERN:
            for q in (0, 1):
                got = getattr(prism, 'kernel' + c)(xp, yp, zp, model[q])
                b = tuple(d['bounds'][q:q + 1, k].copy() for k in range(6))
                cond = oracle.condition('g' + c, xp, yp, zp, b, np.ones((1, 1)), scale=1.0)
                assert_parity(got, d['kernel%s_%d' % (c, q)], cond, floor_fast(1), 'kernel' + c)


def test_overrides_and_skipping(gpu_lib, oracle):
    from fatiando_b200 import prism, mesher
    d = load_golden("skip_override")
    model = prisms_from_golden(d, mesher.Prism)
    xp, yp, zp = d['xp'], d['yp'], d['zp']
    keep = ~d['isnone']
    b = tuple(np.ascontiguousarray(d['bounds'][keep][:, k]) for k in range(6))
    m = int(keep.sum())
    inc, dec = float(d['inc']), float(d['dec'])
    fv = _fvec(d)
    for f in GRAV:
        cond = oracle.condition(f, xp, yp, zp, b, np.full((m, 1), -500.0))
        assert_parity(getattr(prism, f)(xp, yp, zp, model, dens=-500), d[f + '_dens'], cond,
                      floor_fast(m), f + ' dens=')
    pm = d['pmag']
    for f in ['bx', 'by', 'bz']:
        cond = oracle.condition(f, xp, yp, zp, b, np.tile(pm, (m, 1)))
        assert_parity(getattr(prism, f)(xp, yp, zp, model, pmag=pm), d[f + '_pmag'], cond,
                      floor_fast(m), f + ' pmag=')
    props = np.hstack([np.tile(pm, (m, 1)), np.tile(fv, (m, 1))])
    cond = oracle.condition('tf', xp, yp, zp, b, props)
    assert_parity(prism.tf(xp, yp, zp, model, inc, dec, pmag=pm), d['tf_pmag'], cond,
                  floor_fast(m), 'tf pmag=')
    props = np.hstack([np.tile(3.5 * np.array(fv), (m, 1)), np.tile(fv, (m, 1))])
    cond = oracle.condition('tf', xp, yp, zp, b, props)
    assert_parity(prism.tf(xp, yp, zp, model, inc, dec, pmag=3.5), d['tf_pmag_scalar'], cond,
                  floor_fast(m), 'tf pmag scalar')
    # empty and all-skipped models give zeros, like the reference
    assert np.array_equal(prism.gz(xp, yp, zp, []), np.zeros_like(xp))
    assert np.array_equal(prism.gz(xp, yp, zp, [None, model[2]]), np.zeros_like(xp))
    d2 = load_golden("edge_cases")
    sm = [mesher.Prism(-100, 100, -200, 200, 50, 300, {'magnetization': 2.0})]
    cond = _cond(oracle, 'tf', d2, _fvec(d2))
    assert_parity(prism.tf(d2['xp'], d2['yp'], d2['zp'], sm, float(d2['inc']), float(d2['dec'])),
                  d2['tf_scalar_mag'], cond, floor_fast(1), 'scalar magnetization')
    with pytest.raises(TypeError):
        prism.bx(d2['xp'], d2['yp'], d2['zp'], sm)


@pytest.mark.parametrize("case,names", [("prism_mesh", GRAV), ("mesh_magnetic", MAG),
                                        ("prism_relief", ['gz', 'gzz', 'gxy', 'potential'])])
def test_meshes_fused_and_reference_order(gpu_lib, oracle, case, names):
    """Meshes flattened and evaluated in one fused pass, fast and reference-order."""
    from fatiando_b200 import prism, mesher, utils
    d = load_golden(case)
    xp, yp, zp = d['xp'], d['yp'], d['zp']
    fvec = _fvec(d) if 'inc' in d else None
    if case == "prism_relief":
        mesh = mesher.PrismRelief(float(d['relief_ref']), tuple(d['relief_dims']),
                                  (d['relief_x'], d['relief_y'], d['relief_z']))
        mesh.addprop('density', d['relief_density_in'])
        prop = 'density'
    else:
        mesh = mesher.PrismMesh(tuple(d['mesh_bounds']), tuple(int(v) for v in d['mesh_shape']))
        if case == "prism_mesh":
            mesh.addprop('density', d['mesh_density'])
            mesh.mask = [int(v) for v in d['mesh_mask']]
            prop = 'density'
        else:
            mesh.addprop('magnetization', [r for r in d['mesh_magnetization']])
            prop = 'magnetization'
    with prism.Model(mesh, prop) as model:
        fast = model.fields(xp, yp, zp, names, fvec=fvec)
        exact = model.fields(xp, yp, zp, names, fvec=fvec, reference_order=True)
    for f in names:
        cond = _cond(oracle, f, d, fvec)
        assert_parity(fast[f], d[f], cond, floor_fast(mesh.size), '%s/%s fast' % (case, f))
        assert_parity(exact[f], d[f], cond, C_FLOOR_EXACT, '%s/%s exact' % (case, f))
    # single-field calls equal the fused pass exactly (same kernel arithmetic per field)
    one = getattr(prism, names[1])(xp, yp, zp, mesh) if names[1] != 'tf' else None
    if one is not None:
        assert_parity(one, d[names[1]], _cond(oracle, names[1], d, fvec), floor_fast(mesh.size))


def test_cube_faces_symmetries_on_gpu(gpu_lib, oracle):
    """test_prism.py:237-363 (test_around) through the new module, 21x21 faces."""
    from fatiando_b200 import prism, mesher
    d = load_golden("cube_faces")
    model = prisms_from_golden(d, mesher.Prism)
    b = tuple(d['bounds'][:, k].copy() for k in range(6))
    props = d['density'][:, None].copy()
    face = {}
    for g in range(6):
        x, y, z = d['xp_%d' % g], d['yp_%d' % g], d['zp_%d' % g]
        res = prism.fields(x, y, z, model, GRAV)
        for f in GRAV:
            cond = oracle.condition(f, x, y, z, b, props)
            assert_parity(res[f], d['%s_%d' % (f, g)], cond, floor_fast(b[0].size), '%s face %d' % (f, g))
            face[f, g] = res[f]
    aae = np.testing.assert_array_almost_equal
    aae(face['gz', 0], -face['gz', 1], 10)
    aae(face['gx', 2], -face['gx', 3], 10)
    aae(face['gx', 2], -face['gz', 0], 10)
    aae(face['gy', 4], -face['gz', 0], 10)
    aae(face['gxx', 2], face['gzz', 0], 10)
    aae(face['gyy', 4], face['gzz', 0], 10)
    aae(face['gxy', 0], face['gxy', 1], 4)
    aae(face['gxz', 0], -face['gxz', 1], 10)
    aae(face['gyz', 2], face['gyz', 3], 4)
    for g in range(6):
        for a in range(g + 1, 6):
            aae(face['potential', g], face['potential', a], 10)


def test_reference_ffi_shape_accumulates(gpu_lib, oracle):
    """`_prism.<field>(..., res)` adds one prism into res (14 entry points)."""
    from fatiando_b200 import _prism
    d = load_golden("two_prisms")
    xp, yp, zp = d['xp'][::7].copy(), d['yp'][::7].copy(), d['zp'][::7].copy()
    fv = _fvec(d)
    for f in GRAV + MAG:
        res = np.full(xp.size, 0.125)
        ref = np.full(xp.size, 0.125)
        for q in range(2):
            bq = [float(v) for v in d['bounds'][q]]
            if f == 'tf':
                props = list(d['magnetization'][q]) + list(fv)
            elif f in MAG:
                props = list(d['magnetization'][q])
            else:
                props = [float(d['density'][q])]
            getattr(_prism, f)(xp, yp, zp, *bq, *props, res)
            oracle.accumulate(f, xp, yp, zp, *bq, props, ref)
        bounds, pr = golden_model(d, f, fv)
        cond = oracle.condition(f, xp, yp, zp, bounds, pr, scale=1.0) + 0.125
        assert_parity(res, ref, cond, C_FLOOR_EXACT, '_prism.' + f)
        assert not np.array_equal(res, np.full(xp.size, 0.125))


def test_jacobian_against_reference_columns(gpu_lib, oracle):
    from fatiando_b200 import prism, mesher
    d = load_golden("jacobian")
    mesh = mesher.PrismMesh(tuple(d['mesh_bounds']), tuple(int(v) for v in d['mesh_shape']))
    xp, yp, zp = d['xp'], d['yp'], d['zp']
    b = tuple(np.ascontiguousarray(d['bounds'][:, k]) for k in range(6))
    m = b[0].size
    for f in ['gz', 'gzz', 'gxy', 'potential']:
        cond = np.empty((xp.size, m))
        for q in range(m):
            bq = tuple(a[q:q + 1] for a in b)
            cond[:, q] = oracle.condition(f, xp, yp, zp, bq, np.ones((1, 1)))
        jac = prism.jacobian(xp, yp, zp, mesh, f)
        assert jac.shape == (xp.size, m) and jac.flags.c_contiguous
        assert_parity(jac, d['jac_' + f], cond, floor_fast(1), 'jacobian ' + f)
        jt = prism.jacobian(xp, yp, zp, mesh, f, transposed=True)
        assert np.array_equal(jt, jac.T)
        je = prism.jacobian(xp, yp, zp, mesh, f, reference_order=True)
        assert_parity(je, d['jac_' + f], cond, C_FLOOR_EXACT, 'jacobian exact ' + f)
    inc, dec = float(d['inc']), float(d['dec'])
    fv = _fvec(d)
    cond = np.empty((xp.size, m))
    for q in range(m):
        bq = tuple(a[q:q + 1] for a in b)
        cond[:, q] = oracle.condition('tf', xp, yp, zp, bq,
                                      np.array([list(d['jac_tf_pmag']) + list(fv)]))
    jac = prism.jacobian(xp, yp, zp, mesh, 'tf', inc=inc, dec=dec)
    assert_parity(jac, d['jac_tf'], cond, floor_fast(1), 'jacobian tf')
    # masked cells give zero columns, like the reference's per-cell loop
    mesh.mask = [1, 5]
    jm = prism.jacobian(xp, yp, zp, mesh, 'gz')
    assert np.array_equal(jm[:, [1, 5]], np.zeros((xp.size, 2)))
    assert np.array_equal(np.delete(jm, [1, 5], axis=1),
                          prism.jacobian(xp, yp, zp, mesh_cells(mesh), 'gz'))
    # J @ density reproduces the forward field (linearity)
    mesh.mask = []
    dens = np.linspace(-500, 800, m)
    mesh.addprop('density', dens)
    gz = prism.gz(xp, yp, zp, mesh)
    np.testing.assert_allclose(prism.jacobian(xp, yp, zp, mesh, 'gz') @ dens, gz, rtol=1e-9,
                               atol=1e-12 * np.abs(gz).max())


def mesh_cells(mesh):
    return [c for c in mesh if c is not None]


def test_random_models_all_component_sets(gpu_lib, oracle):
    """Seeded random models/points incl. points on faces, edges and vertices."""
    from fatiando_b200 import prism
    from fatiando_b200.flatten import FlatModel
    rs = np.random.RandomState(11)
    n, m = 700, 33
    xp, yp, zp = rs.uniform(-5e3, 5e3, n), rs.uniform(-5e3, 5e3, n), rs.uniform(-800, 1200, n)
    x1 = rs.uniform(-4e3, 3e3, m); x2 = x1 + rs.uniform(10, 1500, m)
    y1 = rs.uniform(-4e3, 3e3, m); y2 = y1 + rs.uniform(10, 1500, m)
    z1 = rs.uniform(0, 1e3, m); z2 = z1 + rs.uniform(10, 1500, m)
    xp[:m], yp[:m], zp[:m] = x1, y1, z2 + 100
    xp[m:2 * m], yp[m:2 * m], zp[m:2 * m] = x2, y1 - 50, z1
    xp[2 * m:3 * m], yp[2 * m:3 * m], zp[2 * m:3 * m] = x1 - 30, y2, z2
    xp[3 * m:4 * m], yp[3 * m:4 * m], zp[3 * m:4 * m] = x1, y2, z1            # vertices
    xp[4 * m:5 * m], yp[4 * m:5 * m], zp[4 * m:5 * m] = 0.5 * (x1 + x2), 0.5 * (y1 + y2), 0.5 * (z1 + z2)
    b = (x1, x2, y1, y2, z1, z2)
    dens = rs.uniform(-1e3, 1e3, m)
    mag = rs.uniform(-2, 2, (m, 3))
    fv = [0.3, 0.5, 0.81]
    flat_g = FlatModel(list(b), dens, None, np.arange(m), m)
    flat_m = FlatModel(list(b), None, mag, np.arange(m), m)
    sets = [['gz'], ['gx', 'gy', 'gz'], GRAV[4:], ['gz'] + GRAV[4:], GRAV[1:], GRAV,
            ['gx', 'gxy', 'gzz'], ['potential', 'gyz']]
    with prism.Model(flat_g, 'density') as model:
        for names in sets:
            res = model.fields(xp, yp, zp, names)
            for f in names:
                ref = oracle.model(f, xp, yp, zp, b, dens[:, None])
                cond = oracle.condition(f, xp, yp, zp, b, dens[:, None])
                assert_parity(res[f], ref, cond, floor_fast(m), 'random %s in %s' % (f, names))
    with prism.Model(flat_m, 'magnetization') as model:
        for names in [['tf'], ['bx'], ['by'], ['bz'], ['bx', 'by', 'bz'], MAG, ['tf', 'bz']]:
            res = model.fields(xp, yp, zp, names, fvec=fv)
            for f in names:
                props = mag if f != 'tf' else np.hstack([mag, np.tile(fv, (m, 1))])
                ref = oracle.model(f, xp, yp, zp, b, props)
                cond = oracle.condition(f, xp, yp, zp, b, props)
                assert_parity(res[f], ref, cond, floor_fast(m), 'random %s in %s' % (f, names))


def test_prism_split_and_many_tiles(gpu_lib, oracle):
    """Few points x many prisms: the prism-split path and ragged tile tails."""
    from fatiando_b200 import prism, mesher
    rs = np.random.RandomState(5)
    mesh = mesher.PrismMesh((0, 5000, 0, 5000, 0, 2000), (7, 23, 31))     # 4991 cells
    mesh.addprop('density', rs.uniform(-1000, 1000, mesh.size))
    mesh.mask = list(range(0, mesh.size, 97))
    xp, yp, zp = rs.uniform(0, 5000, 37), rs.uniform(0, 5000, 37), np.full(37, -150.0)
    from fatiando_b200.flatten import flatten
    flat = flatten(mesh, 'density')
    b = tuple(flat.bounds)
    names = ['gz', 'gxx', 'gxy', 'gxz', 'gyy', 'gyz', 'gzz']
    res = prism.fields(xp, yp, zp, mesh, names)
    for f in names:
        ref = oracle.model(f, xp, yp, zp, b, flat.density[:, None], threads=8)
        cond = oracle.condition(f, xp, yp, zp, b, flat.density[:, None], threads=8)
        assert_parity(res[f], ref, cond, C_FLOOR_FAST, 'split ' + f)
    again = prism.fields(xp, yp, zp, mesh, names)
    for f in names:
        assert np.array_equal(res[f], again[f])        # fixed-order reduction: reproducible


def test_size_independent_properties_at_scale(gpu_lib):
    """Config-sized run (C2: 50x50x20 mesh, 200x200 grid): properties, not an oracle."""
    from fatiando_b200 import prism, mesher, gridder
    rs = np.random.RandomState(0)
    mesh = mesher.PrismMesh((0, 5000, 0, 5000, 0, 2000), (20, 50, 50))
    dens = rs.uniform(-1000, 1000, mesh.size)
    mesh.addprop('density', dens)
    xp, yp, zp = gridder.regular((0, 5000, 0, 5000), (200, 200), z=-150)
    names = ['gxx', 'gxy', 'gxz', 'gyy', 'gyz', 'gzz']
    with prism.Model(mesh, 'density') as model:
        t = model.fields(xp, yp, zp, names)
        lap = t['gxx'] + t['gyy'] + t['gzz']
        scale = np.abs(t['gzz']).max()
        assert np.max(np.abs(lap)) < 1e-9 * scale        # Laplace outside the sources
        # observation split invariance: a shard equals the same rows of the full run --
        # bit for bit when the prisms are summed in list order, to rounding otherwise
        # (the prism split over thread blocks depends on the number of points)
        sl = slice(12345, 23456)
        part = model.fields(xp[sl].copy(), yp[sl].copy(), zp[sl].copy(), names)
        full_lo = model.fields(xp, yp, zp, names, list_order=True)
        part_lo = model.fields(xp[sl].copy(), yp[sl].copy(), zp[sl].copy(), names, list_order=True)
        for f in names:
            assert np.array_equal(part_lo[f], full_lo[f][sl])
            np.testing.assert_allclose(part[f], t[f][sl], rtol=1e-10, atol=1e-12 * scale)
            np.testing.assert_allclose(full_lo[f], t[f], rtol=1e-10, atol=1e-12 * scale)
    # linearity in the property
    mesh.addprop('density', 2.0 * dens)
    t2 = prism.gzz(xp, yp, zp, mesh)
    np.testing.assert_allclose(t2, 2.0 * t['gzz'], rtol=1e-12, atol=1e-12 * scale)


def test_node_sharing_equals_per_cell_and_oracle(gpu_lib, oracle):
    """Regular meshes: the node-sharing kernel (default) against the per-cell kernel and
    the oracle, with masked cells, grid lines in mesh planes and points inside the mesh."""
    from fatiando_b200 import prism, mesher, gridder, utils
    from fatiando_b200.flatten import flatten
    rs = np.random.RandomState(21)
    mesh = mesher.PrismMesh((0, 5000, -300, 4100, 10, 2000), (5, 9, 11))
    mesh.addprop('density', rs.uniform(-1000, 1000, mesh.size))
    mesh.addprop('magnetization', rs.uniform(-2, 2, (mesh.size, 3)))
    mesh.mask = [0, 4, 17, 50, 104, 400]
    grids = [gridder.regular((0, 5000, -300, 4100), (23, 19), z=-150),
             # inside the body, on a node level in z, but off the vertical lines through mesh
             # nodes: ON those lines inside the body the tensor is discontinuous and any two
             # roundings of the coordinates disagree (the reference included)
             gridder.regular((100, 4900, -200, 4000), (12, 10), z=408.0),
             gridder.regular((-3e5, 3e5, -3e5, 3e5), (9, 7), z=-4e4)]
    fv = utils.dircos(30, -15)
    for xp, yp, zp in grids:
        for prop, names in (('density', GRAV), ('magnetization', MAG)):
            flat = flatten(mesh, prop)
            props = flat.density[:, None] if prop == 'density' else flat.magnetization
            with prism.Model(mesh, prop) as model:
                assert model.nodes is not None
                nodes = model.fields(xp, yp, zp, names, fvec=fv)
                cells = model.fields(xp, yp, zp, names, fvec=fv, per_cell=True)
            for f in names:
                pr = props if f != 'tf' else np.hstack([props, np.tile(fv, (props.shape[0], 1))])
                ref = oracle.model(f, xp, yp, zp, flat.bounds, pr)
                cond = oracle.condition(f, xp, yp, zp, flat.bounds, pr)
                assert_parity(nodes[f], ref, cond, C_FLOOR_FAST, 'node sharing ' + f)
                assert_parity(cells[f], ref, cond, C_FLOOR_FAST, 'per cell ' + f)
    # dens= override bypasses the node weights (they belong to the mesh's own densities)
    xp, yp, zp = grids[0]
    flat = flatten(mesh, 'density')
    got = prism.gz(xp, yp, zp, mesh, dens=250.0)
    ref = oracle.model('gz', xp, yp, zp, flat.bounds, np.full((flat.size, 1), 250.0))
    cond = oracle.condition('gz', xp, yp, zp, flat.bounds, np.full((flat.size, 1), 250.0))
    assert_parity(got, ref, cond, C_FLOOR_FAST, 'override on a mesh')


def test_node_sharing_jacobian(gpu_lib, oracle):
    """Jacobian of a whole regular mesh by node sharing against the per-cell kernels and the
    oracle's per-cell columns: several fields, both layouts, point counts that do not fill a
    block, a mesh whose node slabs need fewer points per block, points on mesh planes."""
    from fatiando_b200 import prism, mesher, gridder, utils
    from fatiando_b200.flatten import flatten
    cases = [((0, 5000, -300, 4100, 10, 2000), (5, 9, 11), 77),
             ((0, 4000, 0, 300, 0, 2000), (40, 2, 90), 13)]          # slab 41 x 91: 2 points per block
    fv = utils.dircos(30, -15)
    for bounds, shape, npts in cases:
        mesh = mesher.PrismMesh(bounds, shape)
        flat = flatten(mesh, None, override=True)
        assert flat.nodes is not None and not flat.nodes.weights
        xa, ya, za = gridder.scatter((bounds[0] - 500, bounds[1] + 500, bounds[2] - 500, bounds[3] + 500),
                                     npts, z=-120., seed=5)
        # a few points on mesh planes / above mesh nodes
        xa[:3] = flat.nodes.xn[1]; ya[1:4] = flat.nodes.yn[2]
        with prism.Model(mesh, None, keep_all=True) as model:
            for f in ['gz', 'gxx', 'gxy', 'gyz', 'potential', 'tf', 'bz']:
                pm = list(fv) if f in MAG else None
                jn = model.jacobian(xa, ya, za, f, pmag=pm, fvec=fv)
                jc = model.jacobian(xa, ya, za, f, pmag=pm, fvec=fv, per_cell=True)
                jt = model.jacobian(xa, ya, za, f, pmag=pm, fvec=fv, transposed=True)
                assert np.array_equal(jt.T, jn), f
                unit = [1.0] if f not in MAG else (pm + list(fv) if f == 'tf' else pm)
                ref = oracle.jacobian(f, xa, ya, za, flat.bounds, unit, threads=4)
                # per-entry floor: the eight corner terms of that cell
                for q in range(0, flat.size, max(1, flat.size // 40)):
                    bq = [b[q:q + 1] for b in flat.bounds]
                    cond = oracle.condition(f, xa, ya, za, bq, np.array([unit], dtype=float))
                    assert_parity(jn[:, q], ref[:, q], cond, floor_fast(1), 'node jacobian %s col %d' % (f, q))
                    assert_parity(jc[:, q], ref[:, q], cond, floor_fast(1), 'cell jacobian %s col %d' % (f, q))
                scale = np.abs(ref).max()
                assert np.max(np.abs(jn - ref)) <= 1e-9 * scale, f
    # a masked mesh keeps the per-cell kernels (columns scattered back by prism.jacobian)
    mesh = mesher.PrismMesh((0, 1000, 0, 1000, 0, 500), (2, 3, 4))
    mesh.mask = [5]
    assert flatten(mesh, None, override=True).nodes is None
    j = prism.jacobian(xa, ya, za, mesh, 'gz')
    assert j.shape == (xa.size, mesh.size) and np.all(j[:, 5] == 0.0)


def test_hazard_walls_send_merged_forms_to_the_per_cell_kernels(gpu_lib):
    """A mesh whose y spacing makes the reference round two shared walls differently
    (flatten._mismatched_walls): a point lying IN such a wall must be evaluated per cell --
    forward, Jacobian and adjoint then equal the per-cell kernels bit for bit."""
    from fatiando_b200 import prism, mesher, gridder
    from fatiando_b200.flatten import flatten
    mesh = mesher.PrismMesh((0, 5000, -300, 4100, 10, 2000), (5, 9, 11))
    mesh.addprop('density', np.random.RandomState(2).uniform(-1000, 1000, mesh.size))
    flat = flatten(mesh, 'density')
    hazard_y = flat.nodes.hazard[1]
    assert hazard_y.size >= 2 and flat.nodes.hazard[0].size == 0
    xp, yp, zp = gridder.scatter((0, 5000, -300, 4100), 40, z=-80., seed=6)
    data = np.random.RandomState(3).normal(size=xp.size)
    with prism.Model(mesh, 'density') as fwd, prism.Model(mesh, None, keep_all=True) as geo:
        # no point in a hazard wall: the merged forms run (and differ from per cell in the last bits)
        a = fwd.fields(xp, yp, zp, ['gz', 'gyy'])
        b = fwd.fields(xp, yp, zp, ['gz', 'gyy'], per_cell=True)
        assert not np.array_equal(a['gyy'], b['gyy'])
        ja, jb = geo.jacobian(xp, yp, zp, 'gyy'), geo.jacobian(xp, yp, zp, 'gyy', per_cell=True)
        assert not np.array_equal(ja, jb)
        # one point exactly in a mismatched wall: THAT point is re-evaluated per cell (same bits
        # as a per-cell call), every other point keeps its merged-form value
        yp2 = yp.copy()
        yp2[7] = hazard_y[0]
        others = np.arange(xp.size) != 7
        a2 = fwd.fields(xp, yp2, zp, ['gz', 'gyy'])
        assert fwd.last_fallback_points() == 1
        b2 = fwd.fields(xp[7:8].copy(), yp2[7:8].copy(), zp[7:8].copy(), ['gz', 'gyy'], per_cell=True)
        for f in a2:
            assert a2[f][7] == b2[f][0], f
            assert np.array_equal(a2[f][others], a[f][others]), f
        for transposed in (False, True):
            j2 = geo.jacobian(xp, yp2, zp, 'gyy', transposed=transposed)
            assert geo.last_fallback_points() == 1
            jc = geo.jacobian(xp, yp2, zp, 'gyy', per_cell=True, transposed=transposed)
            if transposed:
                j2, jc = j2.T, jc.T
            assert np.array_equal(j2[7], jc[7])
            assert np.array_equal(j2[others], ja[others])
        # the adjoint sums over the points: the flagged point leaves the node-sharing sum (zero
        # weight) and is added by the per-cell kernels
        adj = geo.adjoint(xp, yp2, zp, data, 'gz')
        assert geo.last_fallback_points() == 1
        rest = geo.adjoint(xp[others].copy(), yp2[others].copy(), zp[others].copy(), data[others].copy(), 'gz')
        assert geo.last_fallback_points() == 0
        one = geo.adjoint(xp[7:8].copy(), yp2[7:8].copy(), zp[7:8].copy(), data[7:8].copy(), 'gz', per_cell=True)
        assert np.array_equal(adj, rest + one)
        cells = geo.adjoint(xp, yp2, zp, data, 'gz', per_cell=True)
        np.testing.assert_allclose(adj, cells, rtol=1e-9, atol=1e-10 * np.abs(cells).max())


def test_relief_surface_form_equals_per_cell_and_oracle(gpu_lib, oracle):
    """PrismRelief on a tight regular grid: surface form (default) against the per-cell kernels
    and the oracle; the thickness-dependent shift case falls back by itself."""
    from fatiando_b200 import prism, mesher, gridder
    from fatiando_b200.flatten import flatten
    rs = np.random.RandomState(8)
    area, shape = (0, 8000, -1000, 5000), (19, 23)
    x, y = gridder.regular(area, shape)
    h = 600 * np.sin(x / 900.) * np.cos(y / 700.) + rs.uniform(-80, 80, x.size)     # crosses ref = 0
    for variable in (False, True):
        relief = mesher.PrismRelief(0, gridder.spacing(area, shape)[::-1], (x, y, -h))
        relief.addprop('density', rs.uniform(2000, 3000, x.size) if variable else 2670. * np.ones(x.size))
        flat = flatten(relief, 'density')
        grids = [gridder.regular(area, (17, 16), z=-1500.),
                 gridder.regular((-4e5, 4e5, -4e5, 4e5), (5, 6), z=-3e4),
                 gridder.scatter(area, 200, z=100., seed=3),
                 (flat.bounds[0][:64] + 0.0, flat.bounds[2][:64] + 0.0, -900. * np.ones(64))]
        with prism.Model(relief, 'density') as model:
            assert model.surface is not None
            for xp, yp, zp in grids:
                surf = model.fields(xp, yp, zp, GRAV)
                cells = model.fields(xp, yp, zp, GRAV, per_cell=True)
                one = model.fields(xp, yp, zp, ['gz'])
                for f in GRAV:
                    ref = oracle.model(f, xp, yp, zp, flat.bounds, flat.density[:, None])
                    cond = oracle.condition(f, xp, yp, zp, flat.bounds, flat.density[:, None])
                    assert_parity(surf[f], ref, cond, C_FLOOR_FAST, 'relief surface ' + f)
                    assert_parity(cells[f], ref, cond, C_FLOOR_FAST, 'relief per cell ' + f)
                assert_parity(one['gz'], oracle.model('gz', xp, yp, zp, flat.bounds, flat.density[:, None]),
                              oracle.condition('gz', xp, yp, zp, flat.bounds, flat.density[:, None]),
                              C_FLOOR_FAST, 'relief surface gz alone')
            # points IN the reference plane on grid lines through the rim nodes: gxz / gyz need a
            # cell's own thickness there -> those points are re-evaluated per cell (bit-identical
            # to a per-cell call), the others keep the surface form
            sf = model.surface
            xp = np.repeat(sf.nodes[0][:2], 3)
            yp = np.repeat(sf.nodes[1][:2], 3) + np.tile([7000., -9000., 12000.], 2)
            zp = np.zeros(6)
            a = model.fields(xp, yp, zp, ['gxz', 'gyz', 'gz'])
            k = model.last_fallback_points()
            assert 1 <= k <= 6
            same = np.ones(6, dtype=bool)
            for i in range(6):
                b = model.fields(xp[i:i + 1].copy(), yp[i:i + 1].copy(), zp[i:i + 1].copy(),
                                 ['gxz', 'gyz', 'gz'], per_cell=True)
                for f in a:
                    same[i] &= a[f][i] == b[f][0]
            for f in a:
                ref = oracle.model(f, xp, yp, zp, flat.bounds, flat.density[:, None])
                cond = oracle.condition(f, xp, yp, zp, flat.bounds, flat.density[:, None])
                assert_parity(a[f], ref, cond, C_FLOOR_FAST, 'relief, points in the reference plane ' + f)
            assert same.sum() >= k
    # the drop-in functions use it too, and an override bypasses it
    xp, yp, zp = gridder.regular(area, (9, 8), z=-1200.)
    got = prism.gz(xp, yp, zp, relief)
    ref = oracle.model('gz', xp, yp, zp, flat.bounds, flat.density[:, None])
    cond = oracle.condition('gz', xp, yp, zp, flat.bounds, flat.density[:, None])
    assert_parity(got, ref, cond, C_FLOOR_FAST, 'prism.gz on a relief')
    got = prism.gzz(xp, yp, zp, relief, dens=300.)
    ref = oracle.model('gzz', xp, yp, zp, flat.bounds, np.full((flat.size, 1), 300.))
    cond = oracle.condition('gzz', xp, yp, zp, flat.bounds, np.full((flat.size, 1), 300.))
    assert_parity(got, ref, cond, C_FLOOR_FAST, 'override on a relief')


def test_relief_at_c5_scale(gpu_lib, oracle):
    """BASELINE config 5 at full model size (1000x1000 PrismRelief, 1e6 prisms): surface form
    and per-cell kernels against the oracle on a sample of the 1e6 points, gz + tensor."""
    import bench
    from fatiando_b200 import prism
    from fatiando_b200.flatten import flatten
    w = bench.make_workload("c5t", 1)
    idx = np.linspace(0, w['xp'].size - 1, 24).astype(int)
    xp, yp, zp = (np.ascontiguousarray(w[k][idx]) for k in ('xp', 'yp', 'zp'))
    flat = flatten(w['model'], 'density')
    assert flat.size == 1000000 and flat.surface is not None
    with prism.Model(flat, 'density') as model:
        surf = model.fields(xp, yp, zp, w['names'])
        cells = model.fields(xp, yp, zp, w['names'], per_cell=True)
    for f in w['names']:
        ref = oracle.model(f, xp, yp, zp, flat.bounds, flat.density[:, None], threads=8)
        cond = oracle.condition(f, xp, yp, zp, flat.bounds, flat.density[:, None], threads=8)
        assert_parity(surf[f], ref, cond, C_FLOOR_FAST, 'C5 surface ' + f)
        assert_parity(cells[f], ref, cond, C_FLOOR_FAST, 'C5 per cell ' + f)


def test_c2_at_full_model_size(gpu_lib, oracle):
    """The benchmarked config itself (BASELINE config 2: six tensor components of the 50x50x20
    variable-density mesh) at its FULL model size on 48 of its 200x200 points, node sharing and
    per cell, against the oracle; the same for the grid whose lines lie in mesh planes."""
    import bench
    from fatiando_b200 import prism
    from fatiando_b200.flatten import flatten
    for name in ("c2", "c2a"):
        w = bench.make_workload(name, 1)
        idx = np.linspace(0, w['xp'].size - 1, 48).astype(int)
        xp, yp, zp = (np.ascontiguousarray(w[k][idx]) for k in ('xp', 'yp', 'zp'))
        flat = flatten(w['model'], 'density')
        with prism.Model(flat, 'density') as model:
            assert model.nodes is not None and model.size == 50000
            nodes = model.fields(xp, yp, zp, w['names'])
            cells = model.fields(xp, yp, zp, w['names'], per_cell=True)
        worst = 0.0
        for f in w['names']:
            ref = oracle.model(f, xp, yp, zp, flat.bounds, flat.density[:, None], threads=8)
            cond = oracle.condition(f, xp, yp, zp, flat.bounds, flat.density[:, None], threads=8)
            worst = max(worst, assert_parity(nodes[f], ref, cond, C_FLOOR_FAST, '%s nodes %s' % (name, f)))
            assert_parity(cells[f], ref, cond, C_FLOOR_FAST, '%s per cell %s' % (name, f))
        assert worst < 0.5          # observed: a few hundredths of the floor


def test_borehole_points_on_lines_through_mesh_nodes(gpu_lib, oracle):
    """Points INSIDE a regular mesh on vertical lines through its nodes (borehole data), some
    of them exactly at nodes.  With walls that the reference rounds the same way from both
    sides, every cell around such a line sees the same zero offsets, safe_atan2(0, 0) = 0 and
    safe_log(0) = 0 apply alike, and the node-sharing sum is the reference's regrouped: under
    the floor, like the per-cell kernels.  (Walls rounded differently are hazard walls: the
    call runs per cell, test_hazard_walls_send_merged_forms_to_the_per_cell_kernels.)"""
    from fatiando_b200 import prism, mesher, utils
    from fatiando_b200.flatten import flatten
    rs = np.random.RandomState(31)
    mesh = mesher.PrismMesh((0, 5000, 0, 4000, 0, 2000), (8, 8, 10))       # 500 / 500 / 250 m, exact
    mesh.addprop('density', rs.uniform(-1000, 1000, mesh.size))
    mesh.addprop('magnetization', rs.uniform(-2, 2, (mesh.size, 3)))
    zs = np.concatenate([np.linspace(-100, 2100, 23), [250.0, 1000.0, 0.0, 2000.0]])
    lines = [(1500.0, 2000.0), (0.0, 0.0), (5000.0, 4000.0), (2500.0, 500.0), (2500.0, 730.0)]
    xp = np.concatenate([np.full(zs.size, x) for x, _ in lines])
    yp = np.concatenate([np.full(zs.size, y) for _, y in lines])
    zp = np.tile(zs, len(lines))
    fv = utils.dircos(30, -15)
    for prop, names in (('density', GRAV), ('magnetization', MAG)):
        flat = flatten(mesh, prop)
        assert flat.nodes is not None and all(h.size == 0 for h in flat.nodes.hazard)
        props = flat.density[:, None] if prop == 'density' else flat.magnetization
        with prism.Model(mesh, prop) as model:
            nodes = model.fields(xp, yp, zp, names, fvec=fv)
            cells = model.fields(xp, yp, zp, names, fvec=fv, per_cell=True)
        for f in names:
            pr = props if f != 'tf' else np.hstack([props, np.tile(fv, (props.shape[0], 1))])
            ref = oracle.model(f, xp, yp, zp, flat.bounds, pr)
            cond = oracle.condition(f, xp, yp, zp, flat.bounds, pr)
            assert_parity(nodes[f], ref, cond, C_FLOOR_FAST, 'borehole nodes ' + f)
            assert_parity(cells[f], ref, cond, C_FLOOR_FAST, 'borehole per cell ' + f)


def test_c3_and_c4_at_full_model_size(gpu_lib, oracle):
    """BASELINE configs 3 and 4 at their full model sizes on samples of their points:
    tf+bx+by+bz of the 1e5-cell magnetized mesh (node sharing and per cell), and rows of the
    250000 x 40000 gz sensitivity matrix (node sharing and per cell), against the oracle."""
    import bench
    from fatiando_b200 import prism
    from fatiando_b200.flatten import flatten
    w = bench.make_workload("c3", 1)
    idx = np.linspace(0, w['xp'].size - 1, 40).astype(int)
    xp, yp, zp = (np.ascontiguousarray(w[k][idx]) for k in ('xp', 'yp', 'zp'))
    flat = flatten(w['model'], 'magnetization', fvec=w['fvec'])
    with prism.Model(flat, 'magnetization') as model:
        assert model.nodes is not None
        nodes = model.fields(xp, yp, zp, w['names'], fvec=w['fvec'])
        cells = model.fields(xp, yp, zp, w['names'], fvec=w['fvec'], per_cell=True)
    for f in w['names']:
        pr = flat.magnetization if f != 'tf' else np.hstack([flat.magnetization,
                                                             np.tile(w['fvec'], (flat.size, 1))])
        ref = oracle.model(f, xp, yp, zp, flat.bounds, pr, threads=8)
        cond = oracle.condition(f, xp, yp, zp, flat.bounds, pr, threads=8)
        assert_parity(nodes[f], ref, cond, C_FLOOR_FAST, 'C3 nodes ' + f)
        assert_parity(cells[f], ref, cond, C_FLOOR_FAST, 'C3 per cell ' + f)
    w = bench.make_workload("c4full", 1)
    idx = np.linspace(0, w['xp'].size - 1, 21).astype(int)
    xp, yp, zp = (np.ascontiguousarray(w[k][idx]) for k in ('xp', 'yp', 'zp'))
    flat = flatten(w['model'], None, override=True)
    with prism.Model(flat, None) as model:
        assert model.nodes is not None and model.size == 40000
        jn = model.jacobian(xp, yp, zp, 'gz')
        jc = model.jacobian(xp, yp, zp, 'gz', per_cell=True)
    ref = oracle.jacobian('gz', xp, yp, zp, flat.bounds, [1.0], threads=8)
    for q in range(0, flat.size, 997):
        bq = [b[q:q + 1] for b in flat.bounds]
        cond = oracle.condition('gz', xp, yp, zp, bq, np.ones((1, 1)))
        assert_parity(jn[:, q], ref[:, q], cond, floor_fast(1), 'C4 node jacobian col %d' % q)
        assert_parity(jc[:, q], ref[:, q], cond, floor_fast(1), 'C4 cell jacobian col %d' % q)
    assert np.max(np.abs(jn - ref)) <= 1e-9 * np.abs(ref).max()


def test_negative_zero_bounds_and_points_on_planes(gpu_lib, oracle):
    """-0.0 bounds (e.g. z = -1*0.0 from a flat relief) with points in those planes:
    the reference tests `x < 0`, false for -0.0; the kernel reads sign bits, so the
    upload canonicalises zeros."""
    from fatiando_b200 import prism
    from fatiando_b200.flatten import FlatModel
    nz = -0.0
    b = [np.array([nz, -50.0, 10.0]), np.array([100.0, nz, 60.0]),
         np.array([nz, -30.0, nz]), np.array([80.0, nz, 40.0]),
         np.array([nz, nz, 5.0]), np.array([60.0, 70.0, 25.0])]
    dens = np.array([1000.0, -700.0, 300.0])
    xp = np.array([0.0, -0.0, 30.0, 0.0, 100.0, 12.0, 0.0, -20.0])
    yp = np.array([5.0, 40.0, 0.0, -0.0, 80.0, -0.0, 0.0, 0.0])
    zp = np.array([0.0, -0.0, -3.0, 0.0, 0.0, -1.0, -7.0, 0.0])
    flat = FlatModel(b, dens, None, np.arange(3), 3)
    names = ['potential', 'gx', 'gy', 'gz', 'gxx', 'gxy', 'gxz', 'gyy', 'gyz', 'gzz']
    with prism.Model(flat, 'density') as model:
        res = model.fields(xp, yp, zp, names)
    for f in names:
        ref = oracle.model(f, xp, yp, zp, b, dens[:, None])
        cond = oracle.condition(f, xp, yp, zp, b, dens[:, None])
        assert_parity(res[f], ref, cond, floor_fast(3), 'negative zero ' + f)


def test_input_contract(gpu_lib):
    from fatiando_b200 import prism, mesher
    model = [mesher.Prism(0, 1, 0, 1, 0, 1, {'density': 1.0})]
    x = np.zeros(4)
    with pytest.raises(ValueError):
        prism.gz(x.astype(np.float32), x.astype(np.float32), x.astype(np.float32), model)
    with pytest.raises(ValueError):
        prism.gz(x.reshape(2, 2), x.reshape(2, 2), x.reshape(2, 2), model)
    with pytest.raises(AttributeError):
        prism.gz([0.0], [0.0], [0.0], model)
    big = np.arange(8.0)
    got = prism.gz(big[::2], big[::2], -big[::2] - 1, model)      # strided views accepted
    want = prism.gz(big[::2].copy(), big[::2].copy(), (-big[::2] - 1).copy(), model)
    assert np.array_equal(got, want)
    assert prism.gz(np.zeros(0), np.zeros(0), np.zeros(0), model).shape == (0,)
    nan = prism.gz(np.array([np.nan]), np.array([0.0]), np.array([-1.0]), model)
    assert np.isnan(nan[0])


def test_multi_gpu_sharding_under_nccl(gpu_lib):
    """One rank per visible GPU (at most 2 here); world 1 still exercises the NCCL path."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    n = min(2, gpu_lib.device_count())
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", "29511",
           os.path.join(ROOT, "tests", "dist_nccl_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "OK" in res.stdout


def test_single_process_multi_device(gpu_lib):
    from fatiando_b200 import prism, mesher, gridder, partition
    rs = np.random.RandomState(3)
    mesh = mesher.PrismMesh((0, 5000, 0, 5000, 0, 2000), (4, 10, 10))
    mesh.addprop('density', rs.uniform(-1000, 1000, mesh.size))
    xp, yp, zp = gridder.regular((0, 5000, 0, 5000), (31, 17), z=-150)
    devices = list(range(gpu_lib.device_count())) * 2          # also >1 shard on one GPU
    got = partition.fields_multi_device(xp, yp, zp, mesh, ['gz', 'gyz'], devices, list_order=True)
    with prism.Model(mesh, 'density') as model:
        want = model.fields(xp, yp, zp, ['gz', 'gyz'], list_order=True)
    for nm in want:
        assert np.array_equal(got[nm], want[nm])


def test_adjoint_is_the_transposed_jacobian(gpu_lib, oracle):
    """J^T d on the fly (imaging.py:116-118): against the reference's stored columns,
    and <J p, d> = <p, J^T d> at a size where J is never formed."""
    from fatiando_b200 import prism, mesher, gridder, utils
    d = load_golden("jacobian")
    mesh = mesher.PrismMesh(tuple(d['mesh_bounds']), tuple(int(v) for v in d['mesh_shape']))
    xp, yp, zp = d['xp'], d['yp'], d['zp']
    rs = np.random.RandomState(9)
    data = rs.normal(size=xp.size)
    for f in ['gz', 'gzz', 'gxy', 'potential']:
        want = d['jac_' + f].T @ data
        got = prism.adjoint(xp, yp, zp, mesh, data, f)
        scale = np.abs(d['jac_' + f]).T @ np.abs(data)
        assert np.all(np.abs(got - want) <= 1e-9 * np.abs(want) + 1e-13 * scale), f
        exact = prism.adjoint(xp, yp, zp, mesh, data, f, reference_order=True)
        assert np.all(np.abs(exact - want) <= 1e-9 * np.abs(want) + 1e-13 * scale), f
    inc, dec = float(d['inc']), float(d['dec'])
    want = d['jac_tf'].T @ data
    got = prism.adjoint(xp, yp, zp, mesh, data, 'tf', inc=inc, dec=dec)
    assert np.allclose(got, want, rtol=1e-9, atol=1e-13 * np.abs(d['jac_tf']).sum(0).max())
    mesh.mask = [1, 5]
    got = prism.adjoint(xp, yp, zp, mesh, data, 'gz')
    assert got.shape == (mesh.size,) and got[1] == 0.0 and got[5] == 0.0
    # dot-product test at C4-like proportions (20 000 points x 4 000 cells)
    mesh = mesher.PrismMesh((0, 10000, 0, 10000, 0, 4000), (10, 20, 20))
    p = rs.uniform(-500, 500, mesh.size)
    mesh.addprop('density', p)
    xp, yp, zp = gridder.regular((0, 10000, 0, 10000), (200, 100), z=-100)
    data = rs.normal(size=xp.size)
    lhs = float(np.dot(prism.gz(xp, yp, zp, mesh), data))
    rhs = float(np.dot(p, prism.adjoint(xp, yp, zp, mesh, data, 'gz')))
    assert abs(lhs - rhs) <= 1e-10 * max(abs(lhs), abs(rhs))
    again = prism.adjoint(xp, yp, zp, mesh, data, 'gz')
    assert np.array_equal(again, prism.adjoint(xp, yp, zp, mesh, data, 'gz'))   # reproducible
    # the whole-mesh adjoint runs by node sharing: against the per-cell kernels and, on a sample
    # of cells, against the oracle's columns with the floor of the summands it adds up
    # (eps * sum_l |d_l| * A_lq, A = the eight corner terms of cell q at point l)
    from fatiando_b200.flatten import flatten
    fv = utils.dircos(30, -15)
    flat = flatten(mesh, None, override=True)
    sample = list(range(0, flat.size, flat.size // 24))
    with prism.Model(mesh, None, keep_all=True) as model:
        assert model.nodes is not None
        for f in ['gz', 'gxy', 'gzz', 'tf']:
            pm = list(fv) if f == 'tf' else None
            unit = [1.0] if pm is None else pm + list(fv)
            a = model.adjoint(xp, yp, zp, data, f, pmag=pm, fvec=fv)
            b = model.adjoint(xp, yp, zp, data, f, pmag=pm, fvec=fv, per_cell=True)
            for q in sample:
                bq = [v[q:q + 1] for v in flat.bounds]
                col = oracle.jacobian(f, xp, yp, zp, bq, unit, threads=4)[:, 0]
                cond = oracle.condition(f, xp, yp, zp, bq, np.array([unit], dtype=float), threads=4)
                want, floor = float(col @ data), C_FLOOR_FAST * EPS * float(cond @ np.abs(data))
                assert abs(a[q] - want) <= 1e-9 * abs(want) + floor, (f, q, 'nodes')
                assert abs(b[q] - want) <= 1e-9 * abs(want) + floor, (f, q, 'cells')
