"""
The sphere family against the REAL reference's vectors
(tests/golden/sphere.npz, from fatiando/gravmag/sphere.py via make_golden_sphere.py; its
'regression' case reproduces the reference's own regression file).

not gpu: the numpy oracle (oracle/sphere_oracle.py) against the golden vectors.
gpu: fatiando_b200.sphere through the C ABI against the golden vectors and the oracle,
mirroring fatiando/gravmag/tests/test_sphere.py.
"""
import numpy as np
import pytest

from conftest import load_golden
from helpers import EPS

FIELDS = 'gz gxx gxy gxz gyy gyz gzz bx by bz tf'.split()
KERNELS = 'xx xy xz yy yz zz'.split()


def _many_model(g, Sphere):
    model = []
    for geo, dens, mag, kind in zip(g['many_geo'], g['many_dens'], g['many_mag'], g['many_kind']):
        if kind == 0:
            model.append(None)
            continue
        p = {}
        if kind in (1, 3):
            p['density'] = float(dens)
        if kind in (2, 3):
            p['magnetization'] = mag.copy()
        model.append(Sphere(*geo, props=p))
    return model


def _oracle_many(g, f, dens=None, pmag=None, terms=False):
    from oracle import sphere_oracle as so
    from fatiando_b200 import utils
    kind = g['many_kind']
    magnetic = f in so.MAGNETIC_FIELDS
    if (pmag if magnetic else dens) is not None:
        keep = kind != 0
        props = np.tile(pmag, (keep.sum(), 1)) if magnetic else np.full(keep.sum(), dens)
    else:
        keep = np.isin(kind, (2, 3) if magnetic else (1, 3))
        props = g['many_mag'][keep] if magnetic else g['many_dens'][keep]
    fv = utils.dircos(float(g['many_inc']), float(g['many_dec']))
    return so.model(f, g['many_x'], g['many_y'], g['many_z'], g['many_geo'][keep], props, fvec=fv,
                    terms=terms)


def test_oracle_reproduces_the_reference():
    from oracle import sphere_oracle as so
    from fatiando_b200 import utils
    g = load_golden('sphere')
    x, y, z = g['reg_x'], g['reg_y'], g['reg_z']
    fv = utils.dircos(float(g['reg_inc']), float(g['reg_dec']))
    geo = g['reg_sphere'][None, :]
    for f in FIELDS:
        props = g['reg_mag'][None, :] if f in so.MAGNETIC_FIELDS else np.array([1.0])
        assert np.array_equal(so.model(f, x, y, z, geo, props, fvec=fv), g['reg_' + f]), f
    for f in FIELDS:
        assert np.array_equal(_oracle_many(g, f), g['many_' + f]), f
        if f in so.MAGNETIC_FIELDS:
            assert np.array_equal(_oracle_many(g, f, pmag=[1.5, -0.5, 2.0]), g['many_pmag_' + f]), f
        else:
            assert np.array_equal(_oracle_many(g, f, dens=-250.), g['many_dens_' + f]), f


def _close(got, want, floor_terms, what):
    tol = 1e-9 * np.abs(want) + 4 * EPS * floor_terms
    assert np.all(np.abs(got - want) <= tol), '%s: max excess %.3g' % (
        what, np.max(np.abs(got - want) / np.maximum(tol, 1e-300)))


@pytest.mark.gpu
def test_sphere_regression_and_kernels(gpu_lib):
    """test_sphere.py::test_sphere_regression / _force_prop / _ignore_none, on the GPU."""
    from fatiando_b200 import sphere, utils
    from fatiando_b200.mesher import Sphere
    g = load_golden('sphere')
    x, y, z = g['reg_x'], g['reg_y'], g['reg_z']
    inc, dec = float(g['reg_inc']), float(g['reg_dec'])
    props = {'density': 1, 'magnetization': g['reg_mag']}
    model = [Sphere(*g['reg_sphere'], props=props)]
    for mdl in (model, [None] * 10 + model):
        for f in FIELDS:
            got = sphere.tf(x, y, z, mdl, inc, dec) if f == 'tf' else getattr(sphere, f)(x, y, z, mdl)
            np.testing.assert_allclose(got, g['reg_' + f], rtol=1e-12, atol=1e-10)
    for k in KERNELS:
        np.testing.assert_allclose(getattr(sphere, 'kernel' + k)(x, y, z, model[0]), g['reg_kernel' + k],
                                   rtol=1e-12, atol=1e-10)
    pmag, dens = -10 * g['reg_mag'], -10
    for f in FIELDS:
        if f == 'tf':
            got = sphere.tf(x, y, z, model, inc, dec, pmag=pmag)
        elif f in ('bx', 'by', 'bz'):
            got = getattr(sphere, f)(x, y, z, model, pmag=pmag)
        else:
            got = getattr(sphere, f)(x, y, z, model, dens=dens)
        np.testing.assert_allclose(got, -10 * g['reg_' + f], rtol=1e-12, atol=1e-10)
    # missing property: ignored (test_sphere.py::test_sphere_ignore_missing_prop)
    extra = [Sphere(0, 0, 300, 100, {'density': 5}), Sphere(10, 0, 300, 100, {'magnetization': [1, 2, 3]})]
    a = sphere.gz(x, y, z, model + extra[1:])
    np.testing.assert_allclose(a, g['reg_gz'], rtol=1e-12, atol=1e-10)
    b = sphere.bz(x, y, z, model + extra[:1])
    np.testing.assert_allclose(b, g['reg_bz'], rtol=1e-12, atol=1e-10)
    with pytest.raises(ValueError):
        sphere.gz(x, y[:-1], z, model)
    with pytest.raises(ValueError):
        sphere.fields(x, y, z, model, ['gx'])


@pytest.mark.gpu
def test_sphere_cloud_fused_overrides_and_jacobian(gpu_lib):
    from fatiando_b200 import sphere, utils
    from fatiando_b200.mesher import Sphere
    from oracle import sphere_oracle as so
    g = load_golden('sphere')
    model = _many_model(g, Sphere)
    x, y, z = g['many_x'], g['many_y'], g['many_z']
    inc, dec = float(g['many_inc']), float(g['many_dec'])
    fused = sphere.fields(x, y, z, model, FIELDS, inc=inc, dec=dec)
    for f in FIELDS:
        one = sphere.tf(x, y, z, model, inc, dec) if f == 'tf' else getattr(sphere, f)(x, y, z, model)
        floor = _oracle_many(g, f, terms=True)
        _close(one, g['many_' + f], floor, f)
        _close(fused[f], g['many_' + f], floor, 'fused ' + f)
        if f in so.MAGNETIC_FIELDS:
            pm = [1.5, -0.5, 2.0]
            got = sphere.tf(x, y, z, model, inc, dec, pmag=pm) if f == 'tf' else getattr(sphere, f)(x, y, z, model, pmag=pm)
            _close(got, g['many_pmag_' + f], _oracle_many(g, f, pmag=pm, terms=True), 'pmag ' + f)
        else:
            got = getattr(sphere, f)(x, y, z, model, dens=-250.)
            _close(got, g['many_dens_' + f], _oracle_many(g, f, dens=-250., terms=True), 'dens ' + f)
    # Jacobian columns = one sphere alone with unit property; J . rho = the field
    keep = np.isin(g['many_kind'], (1, 3))
    jac = sphere.jacobian(x, y, z, model, 'gzz')
    assert jac.shape == (x.size, len(model)) and np.all(jac[:, g['many_kind'] == 0] == 0.0)
    rho = np.where(keep, g['many_dens'], 0.0)
    _close(jac @ rho, g['many_gzz'], np.abs(jac) @ np.abs(rho), 'J rho')
    jt = sphere.jacobian(x, y, z, model, 'tf', inc=inc, dec=dec, transposed=True)
    fv = utils.dircos(inc, dec)
    q = int(np.nonzero(g['many_kind'] != 0)[0][3])
    want = so.model('tf', x, y, z, g['many_geo'][q:q + 1], np.array([fv]), fvec=fv)
    np.testing.assert_allclose(jt[q], want, rtol=1e-11, atol=1e-14 * np.abs(want).max())


@pytest.mark.gpu
def test_equivalent_layer_of_a_point_grid(gpu_lib):
    """
    eqlayer.py:100-111: the Jacobian of a PointGrid layer, column i = gz of point source i with
    unit density -- and the layer's field with its own densities -- from the grid's arrays
    (no per-point Sphere objects), against the numpy oracle.
    """
    from fatiando_b200 import sphere, gridder
    from fatiando_b200.mesher import PointGrid
    from oracle import sphere_oracle as so
    layer = PointGrid((-500., 500., -400., 600.), np.linspace(150., 210., 63), (9, 7))
    rs = np.random.RandomState(11)
    layer.addprop('density', rs.uniform(-300, 300, layer.size))
    x, y, z = gridder.scatter((-600, 600, -600, 600), 300, z=-20., seed=5)
    geo = np.transpose([layer.x, layer.y, layer.z, np.zeros(layer.size) + layer.radius])
    jac = sphere.jacobian(x, y, z, layer, 'gz')
    assert jac.shape == (300, 63)
    for q in (0, 17, 62):
        want = so.model('gz', x, y, z, geo[q:q + 1], np.array([1.0]))
        np.testing.assert_allclose(jac[:, q], want, rtol=1e-12, atol=1e-15 * np.abs(want).max())
    assert np.array_equal(jac, sphere.jacobian(x, y, z, list(layer), 'gz'))
    dens = np.asarray(layer.props['density'])
    got = sphere.gz(x, y, z, layer)
    _close(got, so.model('gz', x, y, z, geo, dens), so.model('gz', x, y, z, geo, dens, terms=True), 'layer gz')
    _close(jac @ dens, got, np.abs(jac) @ np.abs(dens), 'J rho')
