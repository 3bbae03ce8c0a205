"""
Gravity and magnetic fields of homogeneous spheres on the B200 -- a drop-in for
``fatiando.gravmag.sphere`` (reference sphere.py:45-991): :func:`tf`, :func:`bx`,
:func:`by`, :func:`bz`, :func:`gz`, :func:`gxx` ... :func:`gzz` and the unscaled
second derivatives :func:`kernelxx` ... :func:`kernelzz` of the ``1/r`` integral.

Same signatures, units (mGal, Eotvos, nT), skipping rules (``None`` entries and
spheres lacking the property are ignored unless ``dens`` / ``pmag`` is given) and
x-north / y-east / z-down convention as the reference.  The sums run on the GPU
through the prism harness with the sphere evaluator (csrc/sphere.cuh); there is
no CPU path.  :func:`fields` evaluates several components in one pass and
:func:`jacobian` builds the sensitivity matrix the equivalent-layer code asks
for (eqlayer.py:100-111).
"""
import ctypes

import numpy

from . import _lib, utils
from .prism import unit_scale, _as_points, _check_points

GRAVITY_FIELDS = ('gz', 'gxx', 'gxy', 'gxz', 'gyy', 'gyz', 'gzz')
MAGNETIC_FIELDS = ('tf', 'bx', 'by', 'bz')


def _is_point_grid(model):
    """The reference's PointGrid or ours: x, y, z arrays, one radius, per-point property lists."""
    return all(hasattr(model, a) for a in ('area', 'shape', 'radius', 'x', 'y', 'z', 'props', 'dx', 'dy'))


def _flatten(spheres, prop, override):
    """SoA arrays of the spheres a field with property ``prop`` sums (sphere.py:351-357)."""
    if _is_point_grid(spheres) and (override or prop is None or prop in spheres.props):
        # a PointGrid (the equivalent layer, mesh.py:186-388): its arrays ARE the SoA form
        n = spheres.size
        geo = [numpy.ascontiguousarray(a, dtype=numpy.float64) for a in
               (spheres.x, spheres.y, spheres.z, numpy.zeros(n) + spheres.radius)]
        dens = mag = None
        if not override and prop == 'density':
            dens = numpy.ascontiguousarray(spheres.props['density'], dtype=numpy.float64)
        if not override and prop == 'magnetization':
            mag = numpy.ascontiguousarray(spheres.props['magnetization'], dtype=numpy.float64)
        # anything but one value (one vector) per point: per-point objects, as before
        if (dens is None or dens.shape == (n,)) and (mag is None or mag.shape == (n, 3)) \
                and all(a.shape == (n,) for a in geo):
            return geo, dens, mag
    keep = [s for s in spheres
            if s is not None and (override or prop is None or prop in s.props)]
    geo = numpy.array([[s.x, s.y, s.z, s.radius] for s in keep], dtype=numpy.float64).reshape(-1, 4)
    dens = mag = None
    if not override and prop == 'density':
        dens = numpy.array([float(s.props['density']) for s in keep], dtype=numpy.float64)
    if not override and prop == 'magnetization':
        mag = numpy.array([[float(c) for c in s.props['magnetization']] for s in keep],
                          dtype=numpy.float64).reshape(-1, 3)
    return [numpy.ascontiguousarray(geo[:, c]) for c in range(4)], dens, mag


class Model(object):
    """Spheres resident in GPU memory; use as a context manager or call :meth:`close`."""

    def __init__(self, spheres, prop='density', keep_all=False, device=0):
        geo, dens, mag = _flatten(spheres, prop, keep_all)
        self.size = int(geo[0].size)
        self.device = device
        self._handle = ctypes.c_void_p()
        cols = [None, None, None] if mag is None else \
            [numpy.ascontiguousarray(mag[:, c]) for c in range(3)]
        _lib.check(_lib.lib().fb200_model_create_spheres(
            device, self.size, *[_lib.dptr(a) for a in geo], _lib.dptr(dens),
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
        _check_points(xp, yp, zp, 'shape')
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

    def jacobian(self, xp, yp, zp, field='gz', pmag=None, fvec=None, transposed=False, scaled=True):
        xp, yp, zp = _as_points(xp), _as_points(yp), _as_points(zp)
        _check_points(xp, yp, zp, 'shape')
        n, m = xp.size, self.size
        shape = (m, n) if transposed else (n, m)
        out = numpy.zeros(shape, dtype=numpy.float64)
        if n and m:
            unit = numpy.array([1.0]) if field in GRAVITY_FIELDS else \
                numpy.ascontiguousarray(pmag, dtype=numpy.float64)
            fvec_arr = None if fvec is None else numpy.ascontiguousarray(fvec, dtype=numpy.float64)
            _lib.check(_lib.lib().fb200_jacobian(
                self._handle, xp.ctypes.data, yp.ctypes.data, zp.ctypes.data, n,
                _lib.FIELD_ID[field], _lib.dptr(unit), _lib.dptr(fvec_arr),
                unit_scale(field) if scaled else 1.0, out.ctypes.data, shape[1],
                _lib.TRANSPOSED if transposed else 0))
        return out


def fields(xp, yp, zp, spheres, names, dens=None, pmag=None, inc=None, dec=None, device=0):
    """Several components in ONE pass (not in the reference); returns ``{name: array}``."""
    names = list(names)
    unknown = [nm for nm in names if nm not in GRAVITY_FIELDS + MAGNETIC_FIELDS]
    if unknown:
        raise ValueError("spheres offer %s, not %s" % (GRAVITY_FIELDS + MAGNETIC_FIELDS, unknown))
    fvec = None if inc is None else utils.dircos(inc, dec)
    res = {}
    grav = [nm for nm in names if nm in GRAVITY_FIELDS]
    magn = [nm for nm in names if nm in MAGNETIC_FIELDS]
    if grav:
        with Model(spheres, 'density', keep_all=dens is not None, device=device) as model:
            res.update(model.fields(xp, yp, zp, grav, dens=dens))
    if magn:
        if 'tf' in magn and fvec is None:
            raise ValueError("'tf' needs the inducing field direction inc, dec")
        with Model(spheres, 'magnetization', keep_all=pmag is not None, device=device) as model:
            res.update(model.fields(xp, yp, zp, magn, pmag=pmag, fvec=fvec))
    return res


def jacobian(xp, yp, zp, spheres, field='gz', inc=None, dec=None, pmag=None, transposed=False,
             device=0):
    """
    Sensitivity matrix: column q is ``sphere.<field>(x, y, z, [sphere_q], dens=1)`` (or with
    magnetization ``pmag``, default the unit vector along ``inc, dec``), the way the
    equivalent layer builds it from point sources (eqlayer.py:100-111, :145-157).
    ``None`` entries give zero columns.
    """
    fvec = None if inc is None else utils.dircos(inc, dec)
    if field in MAGNETIC_FIELDS:
        if field == 'tf' and fvec is None:
            raise ValueError("'tf' needs the inducing field direction inc, dec")
        if pmag is None:
            if fvec is None:
                raise ValueError("magnetic Jacobians need pmag or inc, dec")
            pmag = fvec
    if _is_point_grid(spheres):
        # the equivalent layer itself (eqlayer.py:108-110 walks it point by point)
        with Model(spheres, None, keep_all=True, device=device) as model:
            return model.jacobian(xp, yp, zp, field, pmag=pmag, fvec=fvec, transposed=transposed)
    spheres = list(spheres)
    kept = [i for i, s in enumerate(spheres) if s is not None]
    with Model(spheres, None, keep_all=True, device=device) as model:
        jac = model.jacobian(xp, yp, zp, field, pmag=pmag, fvec=fvec, transposed=transposed)
    if len(kept) == len(spheres):
        return jac
    full = numpy.zeros((len(spheres), xp.size) if transposed else (xp.size, len(spheres)))
    if transposed:
        full[kept, :] = jac
    else:
        full[:, kept] = jac
    return full


def _gravity(name, xp, yp, zp, spheres, dens):
    if xp.shape != yp.shape or xp.shape != zp.shape:
        raise ValueError("Input arrays xp, yp, and zp must have same shape!")
    return fields(xp, yp, zp, spheres, [name], dens=dens)[name]


def _magnetic(name, xp, yp, zp, spheres, pmag, inc=None, dec=None):
    if xp.shape != yp.shape or xp.shape != zp.shape:
        raise ValueError("Input arrays xp, yp, and zp must have same shape!")
    return fields(xp, yp, zp, spheres, [name], pmag=pmag, inc=inc, dec=dec)[name]


def tf(xp, yp, zp, spheres, inc, dec, pmag=None):
    """Total-field anomaly in nT (sphere.py:45-127).  ``pmag``: [mx, my, mz] override."""
    return _magnetic('tf', xp, yp, zp, spheres, pmag, inc, dec)


def bx(xp, yp, zp, spheres, pmag=None):
    """x component of the magnetic induction in nT (sphere.py:130-189)."""
    return _magnetic('bx', xp, yp, zp, spheres, pmag)


def by(xp, yp, zp, spheres, pmag=None):
    """y component of the magnetic induction in nT (sphere.py:192-251)."""
    return _magnetic('by', xp, yp, zp, spheres, pmag)


def bz(xp, yp, zp, spheres, pmag=None):
    """z component of the magnetic induction in nT (sphere.py:254-313)."""
    return _magnetic('bz', xp, yp, zp, spheres, pmag)


def gz(xp, yp, zp, spheres, dens=None):
    """Vertical gravitational attraction in mGal (sphere.py:316-373)."""
    return _gravity('gz', xp, yp, zp, spheres, dens)


def gxx(xp, yp, zp, spheres, dens=None):
    """xx gravity-gradient component in Eotvos (sphere.py:376-435)."""
    return _gravity('gxx', xp, yp, zp, spheres, dens)


def gxy(xp, yp, zp, spheres, dens=None):
    """xy gravity-gradient component in Eotvos (sphere.py:438-497)."""
    return _gravity('gxy', xp, yp, zp, spheres, dens)


def gxz(xp, yp, zp, spheres, dens=None):
    """xz gravity-gradient component in Eotvos (sphere.py:500-559)."""
    return _gravity('gxz', xp, yp, zp, spheres, dens)


def gyy(xp, yp, zp, spheres, dens=None):
    """yy gravity-gradient component in Eotvos (sphere.py:562-621)."""
    return _gravity('gyy', xp, yp, zp, spheres, dens)


def gyz(xp, yp, zp, spheres, dens=None):
    """yz gravity-gradient component in Eotvos (sphere.py:624-683)."""
    return _gravity('gyz', xp, yp, zp, spheres, dens)


def gzz(xp, yp, zp, spheres, dens=None):
    """zz gravity-gradient component in Eotvos (sphere.py:686-745)."""
    return _gravity('gzz', xp, yp, zp, spheres, dens)


def _kernel(name, xp, yp, zp, sphere):
    # volume * second derivative of 1/r, unscaled (sphere.py:748-991) = the tensor component
    # of this one sphere with unit density and no unit factor
    if xp.shape != yp.shape or xp.shape != zp.shape:
        raise ValueError("Input arrays xp, yp, and zp must have same shape!")
    with Model([sphere], None, keep_all=True) as model:
        return model.fields(xp, yp, zp, [name], dens=1.0, scaled=False)[name]


def kernelxx(xp, yp, zp, sphere):
    """xx second derivative of the sphere's ``1/r`` integral (sphere.py:748-786)."""
    return _kernel('gxx', xp, yp, zp, sphere)


def kernelxy(xp, yp, zp, sphere):
    """xy second derivative (sphere.py:789-827)."""
    return _kernel('gxy', xp, yp, zp, sphere)


def kernelxz(xp, yp, zp, sphere):
    """xz second derivative (sphere.py:830-868)."""
    return _kernel('gxz', xp, yp, zp, sphere)


def kernelyy(xp, yp, zp, sphere):
    """yy second derivative (sphere.py:871-909)."""
    return _kernel('gyy', xp, yp, zp, sphere)


def kernelyz(xp, yp, zp, sphere):
    """yz second derivative (sphere.py:912-950)."""
    return _kernel('gyz', xp, yp, zp, sphere)


def kernelzz(xp, yp, zp, sphere):
    """zz second derivative (sphere.py:953-991)."""
    return _kernel('gzz', xp, yp, zp, sphere)
