"""
The oracle against what pins it (CPU only):
  * tests/golden/*.npz -- outputs of the REAL reference (make_golden.py);
  * oracle/_ref -- the reference's own Cython kernel, compiled unmodified, bit for bit;
  * SURVEY.md App. F known answers and the symmetries of the reference's test_around.
"""
import numpy as np
import pytest

from conftest import load_golden
from helpers import GRAV, MAG, KERN, golden_model

DIRCOS = None


def _fvec(d):
    from fatiando_b200 import utils
    return utils.dircos(float(d['inc']), float(d['dec']))


@pytest.mark.parametrize("case", ["two_prisms", "c1_cookbook", "edge_cases", "skip_override",
                                  "prism_mesh", "mesh_magnetic", "prism_relief"])
def test_oracle_reproduces_reference_fields_bitwise(oracle, case):
    d = load_golden(case)
    fvec = _fvec(d) if 'inc' in d else None
    checked = 0
    for f in GRAV + MAG:
        if f not in d:
            continue
        bounds, props = golden_model(d, f, fvec)
        got = oracle.model(f, d['xp'], d['yp'], d['zp'], bounds, props)
        # same compiler flags, same libm, same operation order: identical bits
        assert np.array_equal(got, d[f]), '%s/%s max diff %g' % (case, f, np.abs(got - d[f]).max())
        checked += 1
    assert checked > 0


def test_oracle_kernels_and_overrides(oracle):
    d = load_golden("two_prisms")
    for c in KERN:
        for q in range(2):
            b = tuple(d['bounds'][q:q + 1, k].copy() for k in range(6))
            got = oracle.model('g' + c, d['xp'], d['yp'], d['zp'], b, np.ones((1, 1)), scale=1.0)
            assert np.array_equal(got, d['kernel%s_%d' % (c, q)])
    d = load_golden("skip_override")
    keep = ~d['isnone']
    b = tuple(np.ascontiguousarray(d['bounds'][keep][:, k]) for k in range(6))
    m = int(keep.sum())
    for f in GRAV:     # dens= applies to every non-None prism (prism.py:132-137)
        got = oracle.model(f, d['xp'], d['yp'], d['zp'], b, np.full((m, 1), -500.0))
        assert np.array_equal(got, d[f + '_dens'])
    fv = _fvec(d)
    for f in ['bx', 'by', 'bz']:
        got = oracle.model(f, d['xp'], d['yp'], d['zp'], b, np.tile(d['pmag'], (m, 1)))
        assert np.array_equal(got, d[f + '_pmag'])
    props = np.hstack([np.tile(d['pmag'], (m, 1)), np.tile(fv, (m, 1))])
    assert np.array_equal(oracle.model('tf', d['xp'], d['yp'], d['zp'], b, props), d['tf_pmag'])
    pm = np.array([3.5 * fv[0], 3.5 * fv[1], 3.5 * fv[2]])
    props = np.hstack([np.tile(pm, (m, 1)), np.tile(fv, (m, 1))])
    assert np.array_equal(oracle.model('tf', d['xp'], d['yp'], d['zp'], b, props),
                          d['tf_pmag_scalar'])


def test_oracle_jacobian_matches_reference_columns(oracle):
    d = load_golden("jacobian")
    b = tuple(np.ascontiguousarray(d['bounds'][:, k]) for k in range(6))
    for f in ['gz', 'gzz', 'gxy', 'potential']:
        jac = oracle.jacobian(f, d['xp'], d['yp'], d['zp'], b, [1.0])
        assert np.array_equal(jac, d['jac_' + f])
    fv = _fvec(d)
    jac = oracle.jacobian('tf', d['xp'], d['yp'], d['zp'], b, list(d['jac_tf_pmag']) + list(fv))
    assert np.array_equal(jac, d['jac_tf'])


def test_oracle_cube_faces_and_symmetries(oracle):
    d = load_golden("cube_faces")
    b = tuple(d['bounds'][:, k].copy() for k in range(6))
    props = d['density'][:, None].copy()
    face = {}
    for g in range(6):
        for f in GRAV:
            got = oracle.model(f, d['xp_%d' % g], d['yp_%d' % g], d['zp_%d' % g], b, props)
            assert np.array_equal(got, d['%s_%d' % (f, g)])
            face[f, g] = got
    # the reference's own consistency checks (test_prism.py:237-363), a sample
    np.testing.assert_array_almost_equal(face['gz', 0], -face['gz', 1], 10)
    np.testing.assert_array_almost_equal(face['gx', 2], -face['gz', 0], 10)
    np.testing.assert_array_almost_equal(face['gxx', 2], face['gzz', 0], 10)
    np.testing.assert_array_almost_equal(face['gxy', 0], face['gxy', 1], 4)   # shift side
    for g in range(6):
        lap = face['gxx', g] + face['gyy', g] + face['gzz', g]
        assert np.max(np.abs(lap)) < 1e-9     # Laplace outside the body


def test_oracle_known_answers(oracle):
    """SURVEY.md App. F (probe of the reference), 10 significant digits."""
    d = load_golden("edge_cases")
    fv = _fvec(d)
    table = {
        'potential': [0.004875855424, 0.004272915887, 0.008662986057, 0.01123701463,
                      0.006624698951, 0.005618507316],
        'gz': [1.335827956, -1.085339352, 4.508508929, None, 0.0, 1.644581841],
        'gxx': [-41.41797638, -31.87710021, -261.5612011, -412.766066, -275.5516532, -156.4149971],
        'gxy': [52.918119, 24.19846643, 0.0, 0.0, 1375.915575, 341.0061561],
        'gzz': [55.84224099, 44.57215476, 356.7804771, -298.0398782, -125.869499, 67.56425411],
        'bx': [-251.5139892, 9.961626686, -668.5967858, -1055.103218, -833.02436, -388.9830906],
        'tf': [-256.2964055, 94.97491504, -309.3478629, -1131.841027, -929.9274285, -304.1784313],
    }
    for f, vals in table.items():
        bounds, props = golden_model(d, f, fv)
        got = oracle.model(f, d['xp'][:6], d['yp'][:6], d['zp'][:6], bounds, props)
        for g, v in zip(got, vals):
            if v is None:
                assert abs(g) < 1e-12
            else:
                assert g == pytest.approx(v, rel=2e-9, abs=1e-12), f
    # scalar magnetization == intensity along F; inverted prism flips sign; flat prism = 0
    bounds, _ = golden_model(d, 'bx')
    props = np.hstack([np.array([[2.0 * fv[0], 2.0 * fv[1], 2.0 * fv[2]]]), np.array([fv])])
    assert np.array_equal(oracle.model('tf', d['xp'], d['yp'], d['zp'], bounds, props),
                          d['tf_scalar_mag'])
    # a zero-thickness prism cancels corner against corner: 0 up to rounding of the
    # running sum (the reference itself leaves ~1e-15 mGal at three of these points)
    assert np.max(np.abs(d['gz_flat'])) < 1e-14
    bflat = tuple(np.array([v]) for v in (-100., 100., -200., 200., 50., 50.))
    assert np.array_equal(oracle.model('gz', d['xp'], d['yp'], d['zp'], bflat, np.array([[1000.]])),
                          d['gz_flat'])
    binv = tuple(np.array([v]) for v in (100., -100., -200., 200., 50., 300.))
    assert np.array_equal(oracle.model('gz', d['xp'], d['yp'], d['zp'], binv, np.array([[1000.]])),
                          d['gz_inverted'])
    # inverted bounds flip the sign (the corner order changes, so only to rounding)
    assert np.allclose(d['gz_inverted'], -d['gz'], rtol=1e-3, atol=1e-12)


def test_oracle_matches_compiled_reference_bitwise(oracle):
    """oracle/_ref is the unmodified _prism.pyx; skip only if it was never built."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    rs = np.random.RandomState(7)
    n, m = 500, 9
    xp, yp, zp = rs.uniform(-5e3, 5e3, n), rs.uniform(-5e3, 5e3, n), rs.uniform(-800, 100, n)
    x1 = rs.uniform(-4e3, 3e3, m); x2 = x1 + rs.uniform(10, 1500, m)
    y1 = rs.uniform(-4e3, 3e3, m); y2 = y1 + rs.uniform(10, 1500, m)
    z1 = rs.uniform(0, 1e3, m); z2 = z1 + rs.uniform(10, 1500, m)
    xp[:m], yp[:m], zp[:m] = x1, y1, z2 + 100          # aligned below: shift active
    xp[m:2 * m], yp[m:2 * m], zp[m:2 * m] = x2, y1 - 50, z1
    xp[2 * m:3 * m], yp[2 * m:3 * m], zp[2 * m:3 * m] = x1 - 30, y2, z2
    b = (x1, x2, y1, y2, z1, z2)
    for f in GRAV + MAG:
        if f == 'tf':
            props = np.hstack([rs.uniform(-2, 2, (m, 3)), np.tile([0.3, 0.5, 0.81], (m, 1))])
        elif f in MAG:
            props = rs.uniform(-2, 2, (m, 3))
        else:
            props = rs.uniform(-1e3, 1e3, (m, 1))
        ref = oracle.ref_model(f, xp, yp, zp, b, props)
        assert np.array_equal(oracle.model(f, xp, yp, zp, b, props), ref), f
        assert np.array_equal(oracle.model(f, xp, yp, zp, b, props, threads=4), ref), f


def test_condition_vector_bounds_the_field(oracle):
    d = load_golden("prism_relief")
    bounds, props = golden_model(d, 'gz')
    a = oracle.condition('gz', d['xp'], d['yp'], d['zp'], bounds, props)
    assert np.all(a >= np.abs(d['gz']))
    assert np.median(a / np.abs(d['gz'])) > 1e3      # far field: severe cancellation
