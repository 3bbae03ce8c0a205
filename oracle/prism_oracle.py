"""
ctypes front-end of the CPU oracle (oracle/prism_oracle.c) and loader of the
compiled reference kernel (oracle/_ref/).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs import this module.  The
product package fatiando_b200 never does.

Parity status: pinned -- see the header of prism_oracle.c and
tests/test_oracle.py.
"""
import ctypes
import glob
import importlib.util
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle_prism.so")

FIELDS = ("potential", "gx", "gy", "gz", "gxx", "gxy", "gxz", "gyy", "gyz",
          "gzz", "tf", "bx", "by", "bz")
FIELD_ID = {name: i for i, name in enumerate(FIELDS)}
GRAVITY_FIELDS = FIELDS[:10]
MAGNETIC_FIELDS = FIELDS[10:]

# fatiando/constants.py:19-31 (values restated; products evaluated in Python
# exactly like prism.py:142,190,334,661 do)
G = 0.00000000006673
SI2EOTVOS = 1000000000.0
SI2MGAL = 100000.0
CM = 10. ** (-7)
T2NT = 10. ** (9)


def unit_scale(field):
    """The factor the reference wrapper multiplies res by (prism.py)."""
    if field == "potential":
        return G                      # prism.py:142
    if field in ("gx", "gy", "gz"):
        return G * SI2MGAL            # prism.py:190,238,286
    if field in GRAVITY_FIELDS:
        return G * SI2EOTVOS          # prism.py:334 ... 598
    return CM * T2NT                  # prism.py:661,707,753,799


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle_prism.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = ctypes.CDLL(LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        sz = ctypes.c_size_t
        d = ctypes.c_double
        L.oracle_prism_accumulate.argtypes = [ctypes.c_int, dp, dp, dp, sz,
                                              d, d, d, d, d, d, dp, dp]
        L.oracle_prism_accumulate.restype = None
        common = [ctypes.c_int, dp, dp, dp, sz, sz, dp, dp, dp, dp, dp, dp, dp,
                  d, ctypes.c_int, dp]
        L.oracle_prism_model.argtypes = common
        L.oracle_prism_model.restype = None
        L.oracle_prism_condition.argtypes = common
        L.oracle_prism_condition.restype = None
        L.oracle_prism_jacobian.argtypes = common + [sz]
        L.oracle_prism_jacobian.restype = None
        L.oracle_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def accumulate(field, xp, yp, zp, x1, x2, y1, y2, z1, z2, props, res):
    """Add one prism into res: the `_prism.<field>` contract (_prism.pyx:72-479)."""
    xp, yp, zp = _c(xp), _c(yp), _c(zp)
    props = _c(np.atleast_1d(props))
    assert res.dtype == np.float64 and res.flags.c_contiguous
    lib().oracle_prism_accumulate(FIELD_ID[field], _p(xp), _p(yp), _p(zp),
                                  xp.size, x1, x2, y1, y2, z1, z2, _p(props),
                                  _p(res))
    return res


def _props_matrix(field, m, props):
    pp = 6 if field == "tf" else (3 if field in MAGNETIC_FIELDS else 1)
    props = _c(props).reshape(m, pp) if m else np.zeros((0, pp))
    return props


def model(field, xp, yp, zp, bounds, props, threads=1, scale=None):
    """
    Field of a whole flattened model.  bounds = (x1, x2, y1, y2, z1, z2) arrays
    of length m, props row-major [m][1|3|6] (tf rows: mx,my,mz,fx,fy,fz).
    """
    xp, yp, zp = _c(xp), _c(yp), _c(zp)
    b = [_c(a) for a in bounds]
    m = b[0].size
    props = _props_matrix(field, m, props)
    res = np.zeros(xp.size)
    if scale is None:
        scale = unit_scale(field)
    lib().oracle_prism_model(FIELD_ID[field], _p(xp), _p(yp), _p(zp), xp.size,
                             m, *[_p(a) for a in b], _p(props), scale, threads,
                             _p(res))
    return res


def condition(field, xp, yp, zp, bounds, props, threads=1, scale=None):
    """A_l of SURVEY.md 8(c): scale * sum |property| * sum |terms|."""
    xp, yp, zp = _c(xp), _c(yp), _c(zp)
    b = [_c(a) for a in bounds]
    m = b[0].size
    props = _props_matrix(field, m, props)
    res = np.zeros(xp.size)
    if scale is None:
        scale = unit_scale(field)
    lib().oracle_prism_condition(FIELD_ID[field], _p(xp), _p(yp), _p(zp),
                                 xp.size, m, *[_p(a) for a in b], _p(props),
                                 scale, threads, _p(res))
    return res


def jacobian(field, xp, yp, zp, bounds, unit_props, threads=1, scale=None):
    """(n, m) C-order sensitivity matrix, column q = field of cell q alone."""
    xp, yp, zp = _c(xp), _c(yp), _c(zp)
    b = [_c(a) for a in bounds]
    m = b[0].size
    unit_props = _c(np.atleast_1d(unit_props))
    jac = np.zeros((xp.size, m))
    if scale is None:
        scale = unit_scale(field)
    lib().oracle_prism_jacobian(FIELD_ID[field], _p(xp), _p(yp), _p(zp),
                                xp.size, m, *[_p(a) for a in b],
                                _p(unit_props), scale, threads, _p(jac), m)
    return jac


def max_threads():
    return lib().oracle_max_threads()


# --------------------------------------------------------------------------
# The reference's own compiled kernel (oracle/_ref/, built by build_ref.py)
# --------------------------------------------------------------------------
_ref = None


def ref_available():
    return bool(glob.glob(os.path.join(HERE, "_ref", "_prism*.so")))


def ref_module():
    """Import the unmodified Cython `_prism` extension from oracle/_ref/."""
    global _ref
    if _ref is None:
        paths = glob.glob(os.path.join(HERE, "_ref", "_prism*.so"))
        if not paths:
            raise ImportError("oracle/_ref is not built (run oracle/build_ref.py "
                              "where /root/reference exists)")
        if not hasattr(np, "float"):
            np.float = float  # _prism.pyx:13 evaluates numpy.float at import
        spec = importlib.util.spec_from_file_location("_prism", paths[0])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ref = mod
    return _ref


def ref_model(field, xp, yp, zp, bounds, props, scale=None):
    """
    Drive the reference kernel the way fatiando/gravmag/prism.py does: zeros,
    one `_prism.<field>` call per prism in list order, one final scale
    (prism.py:129-143 and the 13 sibling wrappers).  props rows as in model().
    """
    mod = ref_module()
    func = getattr(mod, field)
    xp, yp, zp = _c(xp), _c(yp), _c(zp)
    b = [_c(a) for a in bounds]
    m = b[0].size
    props = _props_matrix(field, m, props)
    res = np.zeros(xp.size)
    for q in range(m):
        func(xp, yp, zp, b[0][q], b[1][q], b[2][q], b[3][q], b[4][q], b[5][q],
             *[float(v) for v in props[q]], res)
    res *= unit_scale(field) if scale is None else scale
    return res
