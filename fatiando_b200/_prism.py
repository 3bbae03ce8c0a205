"""
The reference's native-module interface, on the GPU.

``fatiando.gravmag._prism`` (fatiando/gravmag/_prism.pyx:72-479) exports
fourteen functions ``f(xp, yp, zp, x1, x2, y1, y2, z1, z2, <props>, res)`` that
ADD one prism's effect into the caller-owned array ``res`` and return nothing.
This module has the same fourteen names with the same argument meaning, so it
can be bound in place of the Cython extension (``prism.py:92-95``; see
INTEGRATION.md).  Each call evaluates in the reference's operation order with
``res[l]`` as the starting value of the running sum.  It exists for strict
parity tests and for unmodified reference callers; the fast way to use the GPU
is :mod:`fatiando_b200.prism`, which hands the kernel a whole model at once.
"""
import numpy

from . import _lib


def _points(xp, yp, zp, res):
    for a in (xp, yp, zp, res):
        if not isinstance(a, numpy.ndarray):
            raise TypeError("Argument has incorrect type (expected numpy.ndarray, got %s)"
                            % type(a).__name__)
        if a.dtype != numpy.float64:
            raise ValueError("Buffer dtype mismatch, expected 'float64' but got '%s'" % a.dtype)
        if a.ndim != 1:
            raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % a.ndim)
    n = len(xp)
    if len(yp) < n or len(zp) < n or len(res) < n:
        raise ValueError("yp, zp and res must be at least as long as xp")
    return n


def _call(name, xp, yp, zp, scalars, res):
    n = _points(xp, yp, zp, res)
    xc, yc, zc = (numpy.ascontiguousarray(a[:n]) for a in (xp, yp, zp))
    rc = res if res.flags.c_contiguous else numpy.ascontiguousarray(res)
    fn = getattr(_lib.lib(), "fb200_prism_" + name)
    _lib.check(fn(_lib.dptr(xc), _lib.dptr(yc), _lib.dptr(zc), n,
                  *[float(v) for v in scalars], _lib.dptr(rc)))
    if rc is not res:
        res[...] = rc


def _gravity(name):
    def accumulate(xp, yp, zp, x1, x2, y1, y2, z1, z2, density, res):
        _call(name, xp, yp, zp, (x1, x2, y1, y2, z1, z2, density), res)
    accumulate.__name__ = name
    accumulate.__doc__ = ("Add the %s effect of one prism into res "
                          "(_prism.pyx `%s`)." % (name, name))
    return accumulate


potential = _gravity("potential")   # _prism.pyx:454-479
gx = _gravity("gx")                 # :196-221
gy = _gravity("gy")                 # :223-248
gz = _gravity("gz")                 # :250-275
gxx = _gravity("gxx")               # :277-302
gxy = _gravity("gxy")               # :304-334
gxz = _gravity("gxz")               # :336-366
gyy = _gravity("gyy")               # :368-393
gyz = _gravity("gyz")               # :395-425
gzz = _gravity("gzz")               # :427-452


def tf(xp, yp, zp, x1, x2, y1, y2, z1, z2, mx, my, mz, fx, fy, fz, res):
    """Add the total-field anomaly of one prism into res (_prism.pyx:72-104)."""
    _call("tf", xp, yp, zp, (x1, x2, y1, y2, z1, z2, mx, my, mz, fx, fy, fz), res)


def bx(xp, yp, zp, x1, x2, y1, y2, z1, z2, mx, my, mz, res):
    """Add the bx effect of one prism into res (_prism.pyx:106-134)."""
    _call("bx", xp, yp, zp, (x1, x2, y1, y2, z1, z2, mx, my, mz), res)


def by(xp, yp, zp, x1, x2, y1, y2, z1, z2, mx, my, mz, res):
    """Add the by effect of one prism into res (_prism.pyx:136-164)."""
    _call("by", xp, yp, zp, (x1, x2, y1, y2, z1, z2, mx, my, mz), res)


def bz(xp, yp, zp, x1, x2, y1, y2, z1, z2, mx, my, mz, res):
    """Add the bz effect of one prism into res (_prism.pyx:166-194)."""
    _call("bz", xp, yp, zp, (x1, x2, y1, y2, z1, z2, mx, my, mz), res)
