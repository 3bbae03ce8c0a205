"""Shared helpers of the test-suite: fixture unpacking and the parity criterion."""
import numpy as np

EPS = 2.0 ** -52

GRAV = ['potential', 'gx', 'gy', 'gz', 'gxx', 'gxy', 'gxz', 'gyy', 'gyz', 'gzz']
MAG = ['tf', 'bx', 'by', 'bz']
KERN = ['xx', 'xy', 'xz', 'yy', 'yz', 'zz']

# The stated tolerance of this project (DESIGN.md "Parity"):
#   |new - ref| <= RTOL*|ref| + C_FLOOR*eps*A
# A = unit * sum_prisms |property| * sum_corners sum_terms |term| is the sum of the
# absolute values of everything that enters the alternating corner sums, computed by
# the instrumented oracle (oracle_prism_condition).  The floor is the "absolute floor
# near zero-crossings" of the north star; C_FLOOR is 0.1 except for models of
# a handful of prisms, see floor_fast.
RTOL = 1e-9
C_FLOOR_FAST = 0.1        # fused fast path and merged forms, many-prism models
C_FLOOR_EXACT = 0.25      # reference-order path (only libdevice vs glibc differ)
C_FEW = 0.6               # see floor_fast


def floor_fast(nprisms):
    """
    The floor constant c for a model of `nprisms` prisms: max(C_FLOOR_FAST, C_FEW / sqrt(nprisms)).

    A is a sum of 24-56 absolute terms per prism and grows like the number of prisms m; what two
    correct fp64 evaluations of the SAME alternating sum disagree by is rounding noise -- the
    reference itself rounds every term (libm log/atan2 to ~0.5 ulp, the product, and 24+ running
    additions into a partial sum of size ~A/4) -- which grows like sqrt(m).  For a single prism
    the reference's own accumulation noise is ~0.2 eps*A rms (8 corners x 3 terms), so no other
    summation order can stay inside 0.1 eps*A there; from m = 36 prisms on the floor is 0.1.
    Observed maxima are printed by tools/accuracy_report.py and carried in bench.py's JSON line.
    """
    return max(C_FLOOR_FAST, C_FEW / float(max(1, nprisms)) ** 0.5)


def parity_excess(new, ref, cond, c_floor, rtol=RTOL):
    """max over elements of |new-ref| / (rtol*|ref| + c*eps*A); <= 1 passes."""
    new, ref, cond = np.asarray(new), np.asarray(ref), np.asarray(cond)
    tol = rtol * np.abs(ref) + c_floor * EPS * cond
    diff = np.abs(new - ref)
    ok = diff <= tol
    if np.all(ok):
        with np.errstate(divide='ignore', invalid='ignore'):
            ratio = np.where(tol > 0, diff / tol, 0.0)
        return float(np.max(ratio)) if ratio.size else 0.0
    with np.errstate(divide='ignore', invalid='ignore'):
        ratio = np.where(tol > 0, diff / tol, np.where(diff > 0, np.inf, 0.0))
    return float(np.max(ratio))


def assert_parity(new, ref, cond, c_floor, what=''):
    new, ref = np.asarray(new), np.asarray(ref)
    assert new.shape == ref.shape, '%s: shape %s vs %s' % (what, new.shape, ref.shape)
    assert np.all(np.isfinite(new) == np.isfinite(ref)), '%s: non-finite mismatch' % what
    excess = parity_excess(new, ref, cond, c_floor)
    assert excess <= 1.0, '%s: parity excess %.3g (max |diff| %.3g, max |ref| %.3g)' % (
        what, excess, np.max(np.abs(new - ref)), np.max(np.abs(ref)))
    return excess


def golden_model(d, field, fvec=None):
    """(bounds tuple, props matrix) of the prisms a golden case keeps for `field`."""
    isnone = d['isnone']
    if field in MAG:
        has = ~np.isnan(d['magnetization'][:, 0]) & ~isnone
        b = d['bounds'][has]
        props = d['magnetization'][has]
        if field == 'tf':
            props = np.hstack([props, np.tile(np.asarray(fvec, float), (props.shape[0], 1))])
    else:
        has = ~np.isnan(d['density']) & ~isnone
        b = d['bounds'][has]
        props = d['density'][has][:, None]
    return tuple(np.ascontiguousarray(b[:, c]) for c in range(6)), np.ascontiguousarray(props)


def prisms_from_golden(d, Prism):
    """Rebuild the list of Prism / None a golden case was generated from."""
    out = []
    for row, dens, mag, isnone in zip(d['bounds'], d['density'], d['magnetization'], d['isnone']):
        if isnone:
            out.append(None)
            continue
        props = {}
        if not np.isnan(dens):
            props['density'] = float(dens)
        if not np.isnan(mag[0]):
            props['magnetization'] = mag.copy() if not np.isnan(mag[1]) else float(mag[0])
        out.append(Prism(*row, props=props))
    return out
