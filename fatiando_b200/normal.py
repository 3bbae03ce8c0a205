"""
Gauss-Newton products on a sensitivity matrix that stays in GPU memory.

The reference builds ``J`` column by column on the host (eqlayer.py:100-111) and hands it to
``fatiando.inversion``: ``Misfit.hessian`` forms ``2 J^T W J`` (misfit.py:250-256),
``Misfit.gradient`` forms ``-2 J^T W r`` (misfit.py:280-294), ``optimization.linear`` solves
the normal equations (optimization.py:51-97).  Here a ROW BLOCK of ``J`` (this GPU's
observations) is built by the Jacobian kernels straight into HBM and consumed there:

* :meth:`SensitivityBlock.gram`  -- ``factor * J^T diag(w) J`` (fp64 register-tiled SYRK,
  csrc/normal.cu), the Hessian of the data misfit;
* :meth:`SensitivityBlock.tdot`  -- ``factor * J^T (w * r)``, its gradient;
* :meth:`SensitivityBlock.dot`   -- ``J p``, the predicted data of a linear problem.

With several GPUs every rank owns the rows of its observations and ONE all-reduce adds the
ranks' Gram matrices and gradients (:func:`fatiando_b200.partition.gram_row_sharded`).  The
matrix-free alternatives need no ``J`` at all: ``prism.fields`` is ``J p`` for a density ``p``
and ``prism.adjoint`` is ``J^T d``.
"""
import ctypes

import numpy

from . import _lib
from .prism import MAGNETIC_FIELDS, _as_points, _check_points, unit_scale


class DeviceArray(object):
    """A float64 array in GPU memory (``fb200_device_alloc``); ``__cuda_array_interface__``
    lets torch / cupy view it without a copy."""

    def __init__(self, shape, device=0):
        self.shape = tuple(int(v) for v in shape)
        self.device = device
        self.size = int(numpy.prod(self.shape, dtype=numpy.int64))
        ptr = ctypes.c_void_p()
        _lib.check(_lib.lib().fb200_device_alloc(device, max(1, self.size) * 8, ctypes.byref(ptr)))
        self.ptr = ptr.value

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": "<f8", "data": (self.ptr, False), "version": 2,
                "strides": None}

    @classmethod
    def from_host(cls, array, device=0):
        array = numpy.ascontiguousarray(array, dtype=numpy.float64)
        out = cls(array.shape, device)
        _lib.check(_lib.lib().fb200_copy(device, out.ptr, array.ctypes.data, array.nbytes, 0))
        return out

    def to_host(self, out=None):
        if out is None:
            out = _lib.pinned_empty(self.shape) if self.size * 8 >= (64 << 20) else numpy.empty(self.shape)
        if out.shape != self.shape or out.dtype != numpy.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array of shape %s" % (self.shape,))
        _lib.check(_lib.lib().fb200_copy(self.device, out.ctypes.data, self.ptr, self.size * 8, 1))
        return out

    def close(self):
        if self.ptr:
            _lib.check(_lib.lib().fb200_device_free(self.device, self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SensitivityBlock(object):
    """
    Rows ``J[l, :]`` of the sensitivity matrix of ``field`` for the observation points
    ``(xp, yp, zp)`` and the cells of ``model`` (a :class:`fatiando_b200.prism.Model`), built
    on the GPU and kept there.  Shape ``(npoints, ncells)``, C-order.
    """

    def __init__(self, model, xp, yp, zp, field='gz', pmag=None, fvec=None, scaled=True, per_cell=False,
                 grid_reuse=False):
        xp, yp, zp = _as_points(xp), _as_points(yp), _as_points(zp)
        _check_points(xp, yp, zp, 'shape')
        self.device = model.device
        self.field = field
        self.shape = (xp.size, model.size)
        n, m = self.shape
        self.J = DeviceArray((n, m), self.device)
        self.build_ms = 0.0
        self.path = 'default'
        if n and m:
            if field in MAGNETIC_FIELDS:
                if pmag is None:
                    raise ValueError("magnetic Jacobians need the unit magnetization pmag")
                unit = numpy.ascontiguousarray(pmag, dtype=numpy.float64)
            else:
                unit = numpy.array([1.0])
            fvec_arr = None if fvec is None else numpy.ascontiguousarray(fvec, dtype=numpy.float64)
            self.path = 'default'
            if grid_reuse and model.nodes is not None and not per_cell:
                # commensurate level grid: tables + pure data movement (prism.grid_plan)
                from .prism import grid_plan
                plan = grid_plan(model.nodes, xp, yp, zp)
                if plan is not None:
                    rc = _lib.lib().fb200_jacobian_grid(
                        model._handle, plan['ni'], plan['nj'], plan['kx'], plan['ky'], _lib.dptr(plan['xu']),
                        _lib.dptr(plan['yv']), _lib.dptr(numpy.ascontiguousarray(plan['zc'])),
                        _lib.FIELD_ID[field], _lib.dptr(unit), _lib.dptr(fvec_arr),
                        unit_scale(field) if scaled else 1.0, self.J.ptr, m, _lib.DEVICE_POINTERS)
                    if rc == _lib.OK:
                        self.path = 'grid'
                        self.build_ms = model.last_kernel_ms()[0]
                    elif rc != _lib.EUNSUP:
                        _lib.check(rc)
            pts = DeviceArray.from_host(numpy.stack([xp, yp, zp]), self.device) if self.path == 'default' else None
            flags = _lib.DEVICE_POINTERS | (_lib.PER_CELL if per_cell else 0)
            if pts is not None:
                try:
                    _lib.check(_lib.lib().fb200_jacobian(
                        model._handle, pts.ptr, pts.ptr + 8 * n, pts.ptr + 16 * n, n, _lib.FIELD_ID[field],
                        _lib.dptr(unit), _lib.dptr(fvec_arr), unit_scale(field) if scaled else 1.0,
                        self.J.ptr, m, flags))
                    self.build_ms = model.last_kernel_ms()[0]
                finally:
                    pts.close()
        self.last_ms = 0.0

    def close(self):
        self.J.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def to_host(self, out=None):
        return self.J.to_host(out)

    def _vector(self, v, size, what):
        if v is None:
            return None
        v = numpy.ascontiguousarray(v, dtype=numpy.float64)
        if v.shape != (size,):
            raise ValueError("%s must have %d entries" % (what, size))
        return DeviceArray.from_host(v, self.device)

    def dot(self, p):
        """``J @ p``: the predicted data for the parameter vector ``p`` (numpy, ``npoints``)."""
        n, m = self.shape
        dp = self._vector(p, m, "p")
        y = DeviceArray((n,), self.device)
        try:
            _lib.check(_lib.lib().fb200_matvec(self.device, self.J.ptr, n, m, m, dp.ptr, y.ptr))
            return y.to_host()
        finally:
            dp.close()
            y.close()

    def tdot(self, r, weights=None, factor=1.0):
        """``factor * J.T @ (weights * r)`` (numpy, ``ncells``): ``Misfit.gradient`` with
        ``factor = -2 * regul_param`` (misfit.py:280-294)."""
        n, m = self.shape
        dr, dw = self._vector(r, n, "r"), self._vector(weights, n, "weights")
        g = DeviceArray((m,), self.device)
        try:
            _lib.check(_lib.lib().fb200_matvec_t(self.device, self.J.ptr, n, m, m, dr.ptr,
                                                 None if dw is None else dw.ptr, float(factor), g.ptr))
            return g.to_host()
        finally:
            dr.close()
            g.close()
            if dw is not None:
                dw.close()

    def gram_device(self, weights=None, factor=1.0, out=None):
        """``factor * J.T @ diag(weights) @ J`` as a :class:`DeviceArray` ``(ncells, ncells)``."""
        n, m = self.shape
        dw = self._vector(weights, n, "weights")
        G = out if out is not None else DeviceArray((m, m), self.device)
        ms = ctypes.c_double(0)
        try:
            _lib.check(_lib.lib().fb200_gram(self.device, self.J.ptr, n, m, m, None if dw is None else dw.ptr,
                                             float(factor), G.ptr, m, ctypes.byref(ms)))
        finally:
            if dw is not None:
                dw.close()
        self.last_ms = ms.value
        return G

    def gram(self, weights=None, factor=1.0):
        """``factor * J.T @ diag(weights) @ J`` on the host: ``Misfit.hessian`` with
        ``factor = 2 * regul_param`` (misfit.py:250-256)."""
        G = self.gram_device(weights, factor)
        try:
            return G.to_host()
        finally:
            G.close()
