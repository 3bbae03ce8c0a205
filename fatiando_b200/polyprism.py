"""
Gravity and magnetic fields of prisms with polygonal cross-section on the B200 -- a drop-in
for ``fatiando.gravmag.polyprism`` (reference polyprism.py:19-1076, Plouff 1976):
:func:`tf`, :func:`bx`, :func:`by`, :func:`bz`, :func:`gz`, :func:`gxx` ... :func:`gzz` and
the second derivatives :func:`kernelxx` ... :func:`kernelzz`.

Same signatures, units and skipping rules as the reference (``None`` entries and prisms
lacking the property are ignored; only :func:`tf` takes a ``pmag`` override).  Every
polygon edge becomes one row of the device model and is evaluated with the reference's own
expressions (csrc/polyprism.cuh) through the prism harness; there is no CPU path.
:func:`fields` evaluates several components in one pass, sharing the per-edge work.
"""
import ctypes

import numpy

from . import _lib, utils
from .prism import unit_scale, _as_points

GRAVITY_FIELDS = ('gz', 'gxx', 'gxy', 'gxz', 'gyy', 'gyz', 'gzz')
MAGNETIC_FIELDS = ('tf', 'bx', 'by', 'bz')


class Model(object):
    """Polygonal prisms resident in GPU memory (one row per polygon edge)."""

    def __init__(self, prisms, prop='density', keep_all=False, device=0):
        keep = [p for p in prisms if p is not None and (keep_all or prop is None or prop in p.props)]
        self.size = len(keep)
        first = numpy.zeros(self.size + 1, dtype=numpy.int64)
        if keep:
            first[1:] = numpy.cumsum([p.nverts for p in keep])
        cat = lambda parts: numpy.ascontiguousarray(
            numpy.concatenate([numpy.asarray(a, dtype=numpy.float64) for a in parts]) if parts
            else numpy.zeros(0))
        vx, vy = cat([p.x for p in keep]), cat([p.y for p in keep])
        z1 = numpy.array([p.z1 for p in keep], dtype=numpy.float64)
        z2 = numpy.array([p.z2 for p in keep], dtype=numpy.float64)
        dens = mag = None
        if not keep_all and prop == 'density':
            dens = numpy.array([float(p.props['density']) for p in keep], dtype=numpy.float64)
        if not keep_all and prop == 'magnetization':
            mag = numpy.array([[float(c) for c in p.props['magnetization']] for p in keep],
                              dtype=numpy.float64).reshape(-1, 3)
        cols = [None, None, None] if mag is None else \
            [numpy.ascontiguousarray(mag[:, c]) for c in range(3)]
        self._handle = ctypes.c_void_p()
        _lib.check(_lib.lib().fb200_model_create_polyprisms(
            device, self.size, first.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
            _lib.dptr(vx), _lib.dptr(vy), _lib.dptr(z1), _lib.dptr(z2), _lib.dptr(dens),
            *[_lib.dptr(c) for c in cols], ctypes.byref(self._handle)))

    def close(self):
        if self._handle:
            _lib.lib().fb200_model_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def fields(self, xp, yp, zp, names, dens=None, pmag=None, fvec=None, scaled=True):
        xp, yp, zp = _as_points(xp), _as_points(yp), _as_points(zp)
        ids = sorted(set(_lib.FIELD_ID[nm] for nm in names))
        n = xp.size
        out = numpy.zeros((len(ids), n), dtype=numpy.float64)
        if n and self.size:
            mask = 0
            for i in ids:
                mask |= 1 << i
            scale = numpy.array([unit_scale(_lib.FIELDS[i]) if scaled else 1.0 for i in ids])
            dens_arr = None if dens is None else numpy.array([float(dens)])
            pmag_arr = None if pmag is None else numpy.ascontiguousarray(pmag, dtype=numpy.float64)
            fvec_arr = None if fvec is None else numpy.ascontiguousarray(fvec, dtype=numpy.float64)
            _lib.check(_lib.lib().fb200_forward(
                self._handle, xp.ctypes.data, yp.ctypes.data, zp.ctypes.data, n, mask,
                _lib.dptr(dens_arr), _lib.dptr(pmag_arr), _lib.dptr(fvec_arr), _lib.dptr(scale),
                out.ctypes.data, n, 0))
        return {_lib.FIELDS[i]: out[c] for c, i in enumerate(ids)}


def _check(xp, yp, zp):
    # the reference's own test is `xp.shape != yp.shape != zp.shape` (polyprism.py:65), which
    # only fires when BOTH differ; the intent -- equal shapes -- is what is enforced here
    if xp.shape != yp.shape or xp.shape != zp.shape:
        raise ValueError("Input arrays xp, yp, and zp must have same shape!")


def fields(xp, yp, zp, prisms, names, pmag=None, inc=None, dec=None, device=0):
    """Several components in ONE pass (not in the reference); returns ``{name: array}``."""
    names = list(names)
    unknown = [nm for nm in names if nm not in GRAVITY_FIELDS + MAGNETIC_FIELDS]
    if unknown:
        raise ValueError("polygonal prisms offer %s, not %s" % (GRAVITY_FIELDS + MAGNETIC_FIELDS, unknown))
    _check(xp, yp, zp)
    fvec = None if inc is None else utils.dircos(inc, dec)
    prisms = list(prisms)
    res = {}
    grav = [nm for nm in names if nm in GRAVITY_FIELDS]
    magn = [nm for nm in names if nm in MAGNETIC_FIELDS]
    if grav:
        with Model(prisms, 'density', device=device) as model:
            res.update(model.fields(xp, yp, zp, grav))
    if magn:
        if 'tf' in magn and fvec is None:
            raise ValueError("'tf' needs the inducing field direction inc, dec")
        if pmag is not None and magn != ['tf']:
            raise ValueError("only tf takes a pmag override (polyprism.py:19)")
        with Model(prisms, 'magnetization', keep_all=pmag is not None, device=device) as model:
            res.update(model.fields(xp, yp, zp, magn, pmag=pmag, fvec=fvec))
    return res


def tf(xp, yp, zp, prisms, inc, dec, pmag=None):
    """Total-field anomaly in nT (polyprism.py:19-83).  ``pmag``: [mx, my, mz] override."""
    return fields(xp, yp, zp, prisms, ['tf'], pmag=pmag, inc=inc, dec=dec)['tf']


def bx(xp, yp, zp, prisms):
    """x component of the magnetic induction in nT (polyprism.py:86-126)."""
    return fields(xp, yp, zp, prisms, ['bx'])['bx']


def by(xp, yp, zp, prisms):
    """y component of the magnetic induction in nT (polyprism.py:129-169)."""
    return fields(xp, yp, zp, prisms, ['by'])['by']


def bz(xp, yp, zp, prisms):
    """z component of the magnetic induction in nT (polyprism.py:172-212)."""
    return fields(xp, yp, zp, prisms, ['bz'])['bz']


def gz(xp, yp, zp, prisms):
    """Vertical gravitational attraction in mGal (polyprism.py:215-308)."""
    return fields(xp, yp, zp, prisms, ['gz'])['gz']


def gxx(xp, yp, zp, prisms):
    """xx gravity-gradient component in Eotvos (polyprism.py:311-350)."""
    return fields(xp, yp, zp, prisms, ['gxx'])['gxx']


def gxy(xp, yp, zp, prisms):
    """xy gravity-gradient component in Eotvos (polyprism.py:353-392)."""
    return fields(xp, yp, zp, prisms, ['gxy'])['gxy']


def gxz(xp, yp, zp, prisms):
    """xz gravity-gradient component in Eotvos (polyprism.py:395-434)."""
    return fields(xp, yp, zp, prisms, ['gxz'])['gxz']


def gyy(xp, yp, zp, prisms):
    """yy gravity-gradient component in Eotvos (polyprism.py:437-476)."""
    return fields(xp, yp, zp, prisms, ['gyy'])['gyy']


def gyz(xp, yp, zp, prisms):
    """yz gravity-gradient component in Eotvos (polyprism.py:479-518)."""
    return fields(xp, yp, zp, prisms, ['gyz'])['gyz']


def gzz(xp, yp, zp, prisms):
    """zz gravity-gradient component in Eotvos (polyprism.py:521-560)."""
    return fields(xp, yp, zp, prisms, ['gzz'])['gzz']


def _kernel(name, xp, yp, zp, prism):
    _check(xp, yp, zp)
    with Model([prism], None, keep_all=True) as model:
        return model.fields(xp, yp, zp, [name], dens=1.0, scaled=False)[name]


def kernelxx(xp, yp, zp, prism):
    """xx second derivative of the prism's ``1/r`` integral (polyprism.py:563-646)."""
    return _kernel('gxx', xp, yp, zp, prism)


def kernelxy(xp, yp, zp, prism):
    """xy second derivative (polyprism.py:649-733)."""
    return _kernel('gxy', xp, yp, zp, prism)


def kernelxz(xp, yp, zp, prism):
    """xz second derivative (polyprism.py:736-823)."""
    return _kernel('gxz', xp, yp, zp, prism)


def kernelyy(xp, yp, zp, prism):
    """yy second derivative (polyprism.py:826-909)."""
    return _kernel('gyy', xp, yp, zp, prism)


def kernelyz(xp, yp, zp, prism):
    """yz second derivative (polyprism.py:912-996)."""
    return _kernel('gyz', xp, yp, zp, prism)


def kernelzz(xp, yp, zp, prism):
    """zz second derivative (polyprism.py:999-1076)."""
    return _kernel('gzz', xp, yp, zp, prism)
