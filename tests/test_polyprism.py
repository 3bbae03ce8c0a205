"""
The polygonal-prism family against the REAL reference's vectors
(tests/golden/polyprism.npz from fatiando/gravmag/polyprism.py via
make_golden_polyprism.py; its 'reg' case reproduces the reference's regression file, its
'box' case is tests/test_polyprism_vs_prism.py's 4-vertex prism with BOTH references'
results).

not gpu: the numpy oracle against the golden vectors (bit for bit) and the device evaluator
(csrc/polyprism.cuh compiled for the host) against the oracle.
gpu: fatiando_b200.polyprism through the C ABI, mirroring the reference's tests; and the
cross-check rectangular prism == 4-vertex polygonal prism on this package's own kernels.
"""
import ctypes

import numpy as np
import pytest

from conftest import load_golden
from test_hostsim import hostsim, dp, _p, NAMES   # noqa: F401  (fixture)

FIELDS = 'gz gxx gxy gxz gyy gyz gzz bx by bz tf'.split()
KERNELS = 'xx xy xz yy yz zz'.split()


def _many(g):
    """[(x, y, z1, z2)], densities, magnetizations, kinds of the 'many' case."""
    ptr = g['many_ptr']
    polys = [(g['many_vx'][a:b], g['many_vy'][a:b], zz[0], zz[1])
             for a, b, zz in zip(ptr[:-1], ptr[1:], g['many_zz'])]
    return polys, g['many_dens'], g['many_mag'], g['many_kind']


def _oracle_many(g, f, pmag=None):
    from oracle import polyprism_oracle as po
    from fatiando_b200 import utils
    polys, dens, mag, kind = _many(g)
    magnetic = f in po.MAGNETIC_FIELDS
    if pmag is not None:
        keep = [q for q in range(len(polys)) if kind[q] != 0]
        props = [pmag] * len(keep)
    else:
        keep = [q for q in range(len(polys)) if kind[q] in ((2, 3) if magnetic else (1, 3))]
        props = [mag[q] if magnetic else dens[q] for q in keep]
    fv = utils.dircos(float(g['many_inc']), float(g['many_dec']))
    return po.model(f, g['many_x'], g['many_y'], g['many_z'], [polys[q] for q in keep], props, fvec=fv)


def test_oracle_reproduces_the_reference():
    from oracle import polyprism_oracle as po
    from fatiando_b200 import utils
    g = load_golden('polyprism')
    x, y, z = g['reg_x'], g['reg_y'], g['reg_z']
    fv = utils.dircos(float(g['reg_inc']), float(g['reg_dec']))
    prism = (g['reg_verts'][:, 0], g['reg_verts'][:, 1], g['reg_z1z2'][0], g['reg_z1z2'][1])
    for f in FIELDS:
        props = [g['reg_mag']] if f in po.MAGNETIC_FIELDS else [1000]
        assert np.array_equal(po.model(f, x, y, z, [prism], props, fvec=fv), g['reg_' + f]), f
    for k in KERNELS:
        assert np.array_equal(po.kernel(k, x, y, z, *prism), g['reg_kernel' + k]), k
    for f in FIELDS:
        assert np.array_equal(_oracle_many(g, f), g['many_' + f]), f
    assert np.array_equal(_oracle_many(g, 'tf', pmag=[0.5, -1.5, 2.5]), g['many_pmag_tf'])


def _edges(polys, props):
    rows = [[] for _ in range(6)]
    pr = [[] for _ in range(3)]
    for (x, y, z1, z2), p in zip(polys, props):
        n = len(x)
        for k in range(n):
            for r, v in zip(rows, (x[k], x[(k + 1) % n], y[k], y[(k + 1) % n], z1, z2)):
                r.append(v)
            for c, v in zip(pr, np.atleast_1d(p).tolist() + [0.0, 0.0]):
                c.append(v)
    return [np.array(r, dtype=float) for r in rows], [np.array(c, dtype=float) for c in pr]


def _close(got, want, what, rtol=1e-9, scale=None):
    # edge terms cancel against each other (closed polygon): floor = a part in 1e11 of the
    # largest value of the field on the grid
    floor = 1e-11 * (np.abs(want).max() if scale is None else scale)
    bad = np.abs(got - want) > rtol * np.abs(want) + floor
    assert not bad.any(), '%s: max |diff| %.3g at |ref| max %.3g' % (what, np.abs(got - want).max(),
                                                                    np.abs(want).max())


def test_device_evaluator_on_the_host_matches_the_oracle(hostsim):
    """csrc/polyprism.cuh compiled for the host (tests/hostsim) against the numpy oracle."""
    from oracle import polyprism_oracle as po
    from fatiando_b200 import utils
    g = load_golden('polyprism')
    polys, dens, mag, kind = _many(g)
    fv = utils.dircos(float(g['many_inc']), float(g['many_dec']))
    x, y, z = (np.ascontiguousarray(g['many_' + c]) for c in 'xyz')
    for mask, magnetic in ((0x3F8, False), (0x008, False), (0x040, False), (0x3C00, True), (0x0800, True)):
        keep = [q for q in range(len(polys)) if kind[q] in ((2, 3) if magnetic else (1, 3))]
        rows, pr = _edges([polys[q] for q in keep], [mag[q] if magnetic else dens[q] for q in keep])
        out = np.zeros((bin(mask).count('1'), x.size))
        barr = (dp * 6)(*[_p(a) for a in rows])
        parr = (dp * 3)(*[_p(a) for a in pr])
        fvv = np.array(fv, dtype=float)
        rc = hostsim.hostsim_sibling_forward(0, ctypes.c_uint32(mask), _p(x), _p(y), _p(z),
                                             ctypes.c_size_t(x.size), ctypes.c_size_t(rows[0].size),
                                             barr, parr, _p(fvv), _p(out))
        assert rc == 0
        for c, f in enumerate([NAMES[i] for i in range(14) if mask >> i & 1]):
            want = g['many_' + f]
            scale = {'gz': po.G * po.SI2MGAL}.get(f, po.G * po.SI2EOTVOS if f[0] == 'g' else po.CM * po.T2NT)
            _close(out[c] * scale, want, 'hostsim %s in 0x%x' % (f, mask))


@pytest.mark.gpu
def test_polyprism_regression_kernels_and_skipping(gpu_lib):
    """test_polyprism.py::test_polyprism_regression and the skipping rules, on the GPU."""
    from fatiando_b200 import polyprism
    from fatiando_b200.mesher import PolygonalPrism
    g = load_golden('polyprism')
    x, y, z = g['reg_x'], g['reg_y'], g['reg_z']
    inc, dec = float(g['reg_inc']), float(g['reg_dec'])
    props = {'density': 1000, 'magnetization': g['reg_mag']}
    model = [PolygonalPrism(g['reg_verts'], g['reg_z1z2'][0], g['reg_z1z2'][1], props)]
    fused = polyprism.fields(x, y, z, model, FIELDS, inc=inc, dec=dec)
    for mdl in (model, [None] * 3 + model):
        for f in FIELDS:
            got = polyprism.tf(x, y, z, mdl, inc, dec) if f == 'tf' else getattr(polyprism, f)(x, y, z, mdl)
            _close(got, g['reg_' + f], f)
    for f in FIELDS:
        _close(fused[f], g['reg_' + f], 'fused ' + f)
    for k in KERNELS:
        _close(getattr(polyprism, 'kernel' + k)(x, y, z, model[0]), g['reg_kernel' + k], 'kernel' + k)
    with pytest.raises(ValueError):
        polyprism.gz(x, y[:-1], z, model)
    with pytest.raises(ValueError):
        polyprism.fields(x, y, z, model, ['gx'])
    # several polygons with None entries and missing properties; tf with pmag
    polys, dens, mag, kind = _many(g)
    mdl = []
    for (px, py, z1, z2), d, m, kd in zip(polys, dens, mag, kind):
        if kd == 0:
            mdl.append(None)
            continue
        p = {}
        if kd in (1, 3):
            p['density'] = float(d)
        if kd in (2, 3):
            p['magnetization'] = m.copy()
        mdl.append(PolygonalPrism(np.transpose([px, py]), z1, z2, p))
    x, y, z = g['many_x'], g['many_y'], g['many_z']
    inc, dec = float(g['many_inc']), float(g['many_dec'])
    for f in FIELDS:
        got = polyprism.tf(x, y, z, mdl, inc, dec) if f == 'tf' else getattr(polyprism, f)(x, y, z, mdl)
        _close(got, g['many_' + f], 'many ' + f)
    _close(polyprism.tf(x, y, z, mdl, inc, dec, pmag=[0.5, -1.5, 2.5]), g['many_pmag_tf'], 'many pmag tf')


@pytest.mark.gpu
def test_prism_equals_four_vertex_polyprism(gpu_lib):
    """tests/test_polyprism_vs_prism.py: the rectangular-prism kernels of this package against
    its polygonal-prism kernels on the same box, with the reference's own tolerances
    (atol 1e-8 gz, 1e-5 tensor and field components... scaled), and both against the
    reference's values for the two formulations."""
    from fatiando_b200 import polyprism, prism
    from fatiando_b200.mesher import PolygonalPrism, Prism
    g = load_golden('polyprism')
    x, y, z = g['box_x'], g['box_y'], g['box_z']
    inc, dec = float(g['box_inc']), float(g['box_dec'])
    props = {'density': 2670., 'magnetization': g['box_mag']}
    v = g['box_verts']
    pp = [PolygonalPrism(v, g['box_z1z2'][0], g['box_z1z2'][1], props)]
    rp = [Prism(v[:, 0].min(), v[:, 0].max(), v[:, 1].min(), v[:, 1].max(), g['box_z1z2'][0],
                g['box_z1z2'][1], props)]
    for f in FIELDS:
        a = polyprism.tf(x, y, z, pp, inc, dec) if f == 'tf' else getattr(polyprism, f)(x, y, z, pp)
        b = prism.tf(x, y, z, rp, inc, dec) if f == 'tf' else getattr(prism, f)(x, y, z, rp)
        _close(a, g['box_poly_' + f], 'polyprism ' + f)
        np.testing.assert_allclose(b, g['box_prism_' + f], rtol=1e-9, atol=1e-12 * np.abs(b).max())
        # the two formulations agree as well as the reference's own two do
        ref_gap = np.abs(g['box_poly_' + f] - g['box_prism_' + f]).max()
        assert np.abs(a - b).max() <= 2 * ref_gap + 1e-9 * np.abs(b).max(), f
