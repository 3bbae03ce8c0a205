r"""
Potential fields of the 3D right rectangular prism, computed on a B200.

Drop-in for ``fatiando.gravmag.prism`` (reference: fatiando/gravmag/prism.py):
the same twenty functions with the same signatures, conventions and units.

.. note:: All input units are SI.  Output is in conventional units: SI for the
    gravitational potential, mGal for gravity, Eotvos for gravity gradients, nT
    for magnetic fields.

.. note:: The coordinate system is x -> North, y -> East and z -> Down.

**Gravity** (Nagy et al., 2000): :func:`potential`, :func:`gx`, :func:`gy`,
:func:`gz`, :func:`gxx`, :func:`gxy`, :func:`gxz`, :func:`gyy`, :func:`gyz`,
:func:`gzz`.  Like the reference, ``gxy``, ``gxz`` and ``gyz`` move the
computation point slightly when it is aligned with a prism edge on the bottom,
east and north side (prism.py "warning", _prism.pyx:327-330).

**Magnetic** (Bhattacharyya, 1964): :func:`tf`, :func:`bx`, :func:`by`,
:func:`bz`.

**Auxiliary**: :func:`kernelxx` ... :func:`kernelzz`, the second derivatives of
:math:`\int\int\int 1/r`.

**What is different underneath.**  The reference calls a scalar C loop once per
prism and per field.  Here the model is flattened once into SoA arrays
(:mod:`fatiando_b200.flatten`), uploaded to HBM, and a fused CUDA kernel
evaluates *all* requested components of *all* prisms in one pass.  Extras that
the reference does not have, built on the same kernels:

* :func:`fields` -- several components in one pass (shares the per-corner work);
* :func:`jacobian` -- the dense sensitivity matrix, one column per cell;
* :func:`adjoint` -- ``J.T @ data`` on the fly, without storing ``J``;
* :class:`Model` -- upload a model once, evaluate many times.

There is no CPU implementation in this package: without the CUDA library or a
GPU these functions raise.
"""
import ctypes

import numpy

from . import _lib, utils
from .constants import G, SI2EOTVOS, CM, T2NT, SI2MGAL
from .flatten import FlatModel, flatten

GRAVITY_FIELDS = ('potential', 'gx', 'gy', 'gz', 'gxx', 'gxy', 'gxz', 'gyy', 'gyz', 'gzz')
MAGNETIC_FIELDS = ('tf', 'bx', 'by', 'bz')

# Coordinates larger than this are evaluated in reference order (the products
# of the collapsed corner sums of the fast path would leave the double range).
FAST_PATH_MAX_COORD = 1e15

# Host-resident Jacobians of at least this many bytes are returned in page-locked memory.
PINNED_OUTPUT_BYTES = 64 << 20

# Regular meshes are evaluated node by node (each mesh node once, instead of once
# per adjacent cell) when this is True; set to False to force the per-cell kernels.
NODE_SHARING = True


def unit_scale(field):
    """The unit factor the reference multiplies each field by."""
    if field == 'potential':
        return G                    # prism.py:142
    if field in ('gx', 'gy', 'gz'):
        return G * SI2MGAL          # prism.py:190, :238, :286
    if field in GRAVITY_FIELDS:
        return G * SI2EOTVOS        # prism.py:334 ... :598
    if field in MAGNETIC_FIELDS:
        return CM * T2NT            # prism.py:661, :707, :753, :799
    raise ValueError("unknown field '%s'" % field)


def _check_points(xp, yp, zp, what):
    # prism.py:127-128 ("...same length!") and :691-692 ("...same shape!")
    if xp.shape != yp.shape or xp.shape != zp.shape:
        raise ValueError("Input arrays xp, yp, and zp must have same %s!" % what)


def _as_points(a):
    # the reference kernel takes float64, 1-D buffers and nothing else
    # (_prism.pyx:72-77): keep that contract, copy only for strides
    if not isinstance(a, numpy.ndarray):
        raise AttributeError("'%s' object has no attribute 'shape'" % type(a).__name__)
    if a.dtype != numpy.float64:
        raise ValueError("Buffer dtype mismatch, expected 'float64' but got '%s'" % a.dtype)
    if a.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % a.ndim)
    return numpy.ascontiguousarray(a)


def grid_plan(nodes, xp, yp, zp):
    """
    The commensurate-grid plan of ``fb200_jacobian_grid`` for the regular mesh ``nodes``
    (:class:`fatiando_b200.flatten.MeshNodes`) and the points ``xp, yp, zp``, or ``None``.

    Applicable when the points are a regular grid at one level in the reference's order
    (gridder/point_generation.py:89-96: x slowest), the cell size is an integer number of grid
    steps on both axes, and -- the part that makes the result EXACT -- the difference
    ``node - point`` is the same double for every (node a, point i) with the same
    ``a*k - i`` (true e.g. for coordinates that are whole metres; checked here for every pair).
    Returns ``dict(ni, nj, kx, ky, xu, yv, zc)``.
    """
    n = xp.size
    if n < 4 or nodes is None or any(h.size for h in nodes.hazard):
        return None
    if not numpy.all(zp == zp[0]):
        return None
    change = numpy.flatnonzero(xp[1:] != xp[:-1])
    nj = n if change.size == 0 else int(change[0]) + 1
    if nj < 2 or n % nj or n // nj < 2:
        return None
    ni = n // nj
    xs, ys = xp[::nj], yp[:nj]
    if not (numpy.array_equal(xp, numpy.repeat(xs, nj)) and numpy.array_equal(yp, numpy.tile(ys, ni))):
        return None

    def axis(coord, pts):
        cells = coord.size - 1
        step, cell = pts[1] - pts[0], coord[1] - coord[0]
        if not (step > 0 and cell > 0):
            return None
        k = int(round(cell / step))
        if k < 1 or abs(k * step - cell) > 1e-9 * cell:
            return None
        diff = coord[:, None] - pts[None, :]
        u = numpy.arange(cells + 1)[:, None] * k - numpy.arange(pts.size)[None, :] + (pts.size - 1)
        table = numpy.ones(cells * k + pts.size)
        table[u.ravel()] = diff.ravel()
        if not numpy.array_equal(table[u], diff):       # same u, different double somewhere
            return None
        return k, table
    ax, ay = axis(nodes.xn, xs), axis(nodes.yn, ys)
    if ax is None or ay is None:
        return None
    return dict(ni=ni, nj=nj, kx=ax[0], ky=ay[0], xu=ax[1], yv=ay[1], zc=nodes.zn - zp[0])


def _max_abs(*arrays):
    return max([float(numpy.max(numpy.abs(a))) for a in arrays if a.size] + [0.0])


class Model(object):
    """
    A prism model resident in GPU memory ("uploaded once").

    * prisms : list of prisms, ``PrismMesh``, ``PrismRelief`` (this package's
      or the reference's) or a :class:`~fatiando_b200.flatten.FlatModel`
    * prop : ``'density'``, ``'magnetization'`` or ``None``; which physical
      property the model carries.  With ``None`` only overrides
      (``dens=`` / ``pmag=``) and Jacobians can be evaluated.
    * keep_all : keep cells lacking ``prop`` (what an override needs,
      prism.py:132-137)
    * fvec : inducing-field direction, to expand scalar magnetizations
    * device : CUDA device index

    * node_sharing : for a regular ``PrismMesh`` also upload the node-sharing
      form (one kernel evaluation per mesh node instead of eight per cell,
      csrc/prism_mesh.cuh), for a ``PrismRelief`` on a tight regular grid its
      surface form (top faces + merged reference-plane nodes,
      csrc/prism_face.cuh); forward fields then use it unless an override or
      ``per_cell=True`` is given.  Same tolerance as the per-cell path.

    Use as a context manager or call :meth:`close`.
    """

    def __init__(self, prisms, prop='density', keep_all=False, fvec=None, device=0,
                 node_sharing=True):
        flat = prisms if isinstance(prisms, FlatModel) else \
            flatten(prisms, prop, override=keep_all, fvec=fvec)
        self.flat = flat
        self.device = device
        self.size = flat.size
        self.max_coord, self.min_edge = flat.extent()
        self._handle = ctypes.c_void_p()
        lib = _lib.lib()
        mag = flat.magnetization
        cols = [None, None, None] if mag is None else \
            [numpy.ascontiguousarray(mag[:, c]) for c in range(3)]
        _lib.check(lib.fb200_model_create(
            device, flat.size, *[_lib.dptr(b) for b in flat.bounds],
            _lib.dptr(flat.density), *[_lib.dptr(c) for c in cols],
            ctypes.byref(self._handle)))
        self.nodes = flat.nodes if (node_sharing and NODE_SHARING) else None
        self.surface = flat.surface if (node_sharing and NODE_SHARING) else None
        if self.surface is not None:
            sf = self.surface
            _lib.check(lib.fb200_model_attach_surface(
                self._handle, sf.faces[0].size, *[_lib.dptr(a) for a in sf.faces],
                sf.nodes[0].size, *[_lib.dptr(a) for a in sf.nodes], _lib.dptr(sf.cell_size)))
        merged = self.nodes if self.nodes is not None else self.surface
        if merged is not None and any(h.size for h in merged.hazard):
            hz = [numpy.ascontiguousarray(h, dtype=numpy.float64) for h in merged.hazard]
            _lib.check(lib.fb200_model_set_hazard_planes(
                self._handle, hz[0].size, _lib.dptr(hz[0]), hz[1].size, _lib.dptr(hz[1]),
                hz[2].size, _lib.dptr(hz[2])))
        if self.nodes is not None:
            nd = self.nodes
            nz, ny, nx = nd.shape
            w = nd.weights
            wd = w[0] if (prop == 'density' and w) else None
            wm = w if (prop == 'magnetization' and w) else [None, None, None]
            _lib.check(lib.fb200_model_attach_mesh(
                self._handle, nx, ny, nz, _lib.dptr(nd.xn), _lib.dptr(nd.yn), _lib.dptr(nd.zn),
                _lib.dptr(nd.cell_size), _lib.dptr(wd), *[_lib.dptr(c) for c in wm]))

    # -- lifetime ----------------------------------------------------------
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

    # -- evaluation --------------------------------------------------------
    def _flags(self, xp, yp, zp, reference_order):
        big = max(self.max_coord, _max_abs(xp, yp, zp))
        if reference_order or not (big <= FAST_PATH_MAX_COORD):
            return _lib.REFERENCE_ORDER
        return 0

    def fields(self, xp, yp, zp, names, dens=None, pmag=None, fvec=None,
               reference_order=False, scaled=True, list_order=False, per_cell=False):
        """
        Evaluate several components in one pass.

        Returns a dict ``{name: array}``.  ``dens`` (float) / ``pmag``
        (3-vector) replace the per-prism property; ``fvec`` is required for
        ``'tf'``.  ``scaled=False`` returns SI values without the unit factor.
        ``list_order=True`` forbids splitting the prism list over thread blocks:
        every point then sums the prisms strictly in list order, so any
        observation shard of a run reproduces the same bits as the full run.
        ``per_cell=True`` bypasses the node-sharing form of a regular mesh.
        """
        xp, yp, zp = _as_points(xp), _as_points(yp), _as_points(zp)
        _check_points(xp, yp, zp, 'shape')
        names = list(names)
        ids = sorted(set(_lib.FIELD_ID[nm] for nm in names))
        mask = 0
        for i in ids:
            mask |= 1 << i
        n = xp.size
        out = numpy.zeros((len(ids), n), dtype=numpy.float64)
        if n and self.size:
            scale = numpy.array([unit_scale(_lib.FIELDS[i]) if scaled else 1.0 for i in ids])
            dens_arr = None if dens is None else numpy.array([dens], dtype=numpy.float64)
            pmag_arr = None if pmag is None else numpy.ascontiguousarray(pmag, dtype=numpy.float64)
            fvec_arr = None if fvec is None else numpy.ascontiguousarray(fvec, dtype=numpy.float64)
            _lib.check(_lib.lib().fb200_forward(
                self._handle, xp.ctypes.data, yp.ctypes.data, zp.ctypes.data, n, mask,
                _lib.dptr(dens_arr), _lib.dptr(pmag_arr), _lib.dptr(fvec_arr),
                _lib.dptr(scale), out.ctypes.data, n,
                self._flags(xp, yp, zp, reference_order)
                | (_lib.NO_PRISM_SPLIT if list_order else 0)
                | (_lib.PER_CELL if per_cell else 0)))
        return {_lib.FIELDS[i]: out[c] for c, i in enumerate(ids)}

    def jacobian(self, xp, yp, zp, field='gz', pmag=None, fvec=None, transposed=False,
                 reference_order=False, scaled=True, out=None, per_cell=False, grid_reuse=False):
        """
        Dense sensitivity matrix of ``field`` for unit physical property.

        ``J[l, q]`` = effect at point l of prism q alone with density 1
        (gravity) or magnetization ``pmag`` (magnetic): column q is what the
        reference's ``prism.<field>(x, y, z, [cell_q], dens=1)`` returns
        (eqlayer.py:108-111).  C-order ``(npoints, nprisms)`` float64, or
        ``(nprisms, npoints)`` with ``transposed=True`` (imaging.py:116-117).

        ``grid_reuse=True`` (opt-in): for a whole regular mesh and points that form a level grid
        commensurate with it (:func:`grid_plan`) the per-corner kernel is evaluated once per
        distinct offset and the matrix is written by pure data movement -- bit-identical to
        the default result, bound by the write of ``J``.  Silently the default path otherwise;
        ``last_jacobian_path`` tells which one ran.
        """
        xp, yp, zp = _as_points(xp), _as_points(yp), _as_points(zp)
        _check_points(xp, yp, zp, 'shape')
        fid = _lib.FIELD_ID[field]
        n, m = xp.size, self.size
        shape = (m, n) if transposed else (n, m)
        if out is None:
            out = None
            if n and m and n * m * 8 >= PINNED_OUTPUT_BYTES:
                # large matrices land in page-locked memory (pooled, _lib.pinned_empty): the row
                # blocks then cross PCIe at DMA speed while the next block is being computed
                try:
                    out = _lib.pinned_empty(shape)
                except RuntimeError:
                    out = None
            if out is None:
                out = numpy.zeros(shape, dtype=numpy.float64)
        if out.shape != shape or out.dtype != numpy.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array of shape %s" % (shape,))
        if n and m:
            if field in MAGNETIC_FIELDS:
                if pmag is None:
                    raise ValueError("magnetic Jacobians need the unit magnetization pmag")
                unit = numpy.ascontiguousarray(pmag, dtype=numpy.float64)
            else:
                unit = numpy.array([1.0])
            fvec_arr = None if fvec is None else numpy.ascontiguousarray(fvec, dtype=numpy.float64)
            flags = self._flags(xp, yp, zp, reference_order)
            self.last_jacobian_path = 'default'
            if grid_reuse and self.nodes is not None and not (transposed or per_cell or reference_order) \
                    and not (flags & _lib.REFERENCE_ORDER):
                plan = grid_plan(self.nodes, xp, yp, zp)
                if plan is not None:
                    rc = _lib.lib().fb200_jacobian_grid(
                        self._handle, plan['ni'], plan['nj'], plan['kx'], plan['ky'], _lib.dptr(plan['xu']),
                        _lib.dptr(plan['yv']), _lib.dptr(numpy.ascontiguousarray(plan['zc'])), fid,
                        _lib.dptr(unit), _lib.dptr(fvec_arr), unit_scale(field) if scaled else 1.0,
                        out.ctypes.data, shape[1], 0)
                    if rc == _lib.OK:
                        self.last_jacobian_path = 'grid'
                        return out
                    if rc != _lib.EUNSUP:
                        _lib.check(rc)
            if transposed:
                flags |= _lib.TRANSPOSED
            if per_cell:
                flags |= _lib.PER_CELL
            _lib.check(_lib.lib().fb200_jacobian(
                self._handle, xp.ctypes.data, yp.ctypes.data, zp.ctypes.data, n, fid,
                _lib.dptr(unit), _lib.dptr(fvec_arr), unit_scale(field) if scaled else 1.0,
                out.ctypes.data, shape[1], flags))
        return out

    def adjoint(self, xp, yp, zp, data, field='gz', pmag=None, fvec=None,
                reference_order=False, scaled=True, per_cell=False):
        """
        ``J.T @ data`` without building ``J``: for every prism q, the sum over
        the points of ``data[l]`` times the unit-property effect of q at l
        (what ``imaging.migrate`` computes layer by layer, imaging.py:116-118).
        Returns a float64 array with one entry per prism of the model.
        """
        xp, yp, zp = _as_points(xp), _as_points(yp), _as_points(zp)
        data = _as_points(data)
        _check_points(xp, yp, zp, 'shape')
        if data.shape != xp.shape:
            raise ValueError("data must have one value per computation point")
        fid = _lib.FIELD_ID[field]
        out = numpy.zeros(self.size, dtype=numpy.float64)
        if self.size:
            if field in MAGNETIC_FIELDS:
                if pmag is None:
                    raise ValueError("magnetic adjoints need the unit magnetization pmag")
                unit = numpy.ascontiguousarray(pmag, dtype=numpy.float64)
            else:
                unit = numpy.array([1.0])
            fvec_arr = None if fvec is None else numpy.ascontiguousarray(fvec, dtype=numpy.float64)
            _lib.check(_lib.lib().fb200_adjoint(
                self._handle, xp.ctypes.data, yp.ctypes.data, zp.ctypes.data, xp.size, fid,
                _lib.dptr(unit), _lib.dptr(fvec_arr), unit_scale(field) if scaled else 1.0,
                data.ctypes.data, out.ctypes.data,
                self._flags(xp, yp, zp, reference_order) | (_lib.PER_CELL if per_cell else 0)))
        return out

    def last_kernel_ms(self):
        """(device milliseconds, kernel launches) of the last evaluation."""
        ms, launches = ctypes.c_double(0), ctypes.c_int(0)
        _lib.check(_lib.lib().fb200_last_kernel_ms(self._handle, ctypes.byref(ms),
                                                   ctypes.byref(launches)))
        return ms.value, launches.value

    def last_fallback_points(self):
        """Points of the last evaluation that the per-cell kernels re-evaluated (hazard walls,
        thickness-dependent shift of a merged relief node)."""
        k = ctypes.c_int(0)
        _lib.check(_lib.lib().fb200_last_fallback_points(self._handle, ctypes.byref(k)))
        return k.value


# --------------------------------------------------------------------------
# one-shot helpers behind the drop-in functions
# --------------------------------------------------------------------------
def _gravity(names, xp, yp, zp, prisms, dens, what='length', scaled=True, **kw):
    _check_points(xp, yp, zp, what)
    with Model(prisms, 'density', keep_all=dens is not None, device=kw.pop('device', 0)) as model:
        return model.fields(xp, yp, zp, names, dens=dens, scaled=scaled, **kw)


def _pmag_vector(pmag, fvec):
    # prism.py:641-645: a scalar pmag is an intensity along the inducing field
    if pmag is None:
        return None
    if isinstance(pmag, (float, int)):
        if fvec is None:
            raise TypeError("pmag must be a 3-component vector")
        return [pmag * fvec[0], pmag * fvec[1], pmag * fvec[2]]
    mx, my, mz = pmag
    return [mx, my, mz]


def _magnetic(names, xp, yp, zp, prisms, pmag, fvec, what, **kw):
    _check_points(xp, yp, zp, what)
    pm = _pmag_vector(pmag, fvec)
    with Model(prisms, 'magnetization', keep_all=pmag is not None, fvec=fvec,
               device=kw.pop('device', 0)) as model:
        return model.fields(xp, yp, zp, names, pmag=pm, fvec=fvec, **kw)


def fields(xp, yp, zp, prisms, names, dens=None, pmag=None, inc=None, dec=None, **kw):
    """
    Several field components in ONE pass over the model (not in the reference).

    * names : any of the fourteen field names; gravity and magnetic names may
      be mixed (they run as two kernels)
    * dens, pmag : property overrides as in the single-field functions
    * inc, dec : inducing-field direction, needed for ``'tf'``

    Returns a dict ``{name: array}`` in the reference's units.
    """
    names = list(names)
    fvec = None if inc is None else utils.dircos(inc, dec)
    res = {}
    grav = [nm for nm in names if nm in GRAVITY_FIELDS]
    magn = [nm for nm in names if nm in MAGNETIC_FIELDS]
    unknown = [nm for nm in names if nm not in GRAVITY_FIELDS + MAGNETIC_FIELDS]
    if unknown:
        raise ValueError("unknown field(s) %s" % unknown)
    if grav:
        res.update(_gravity(grav, xp, yp, zp, prisms, dens, 'shape', **kw))
    if magn:
        if 'tf' in magn and fvec is None:
            raise ValueError("'tf' needs the inducing field direction inc, dec")
        res.update(_magnetic(magn, xp, yp, zp, prisms, pmag, fvec, 'shape', **kw))
    return res


def jacobian(xp, yp, zp, prisms, field='gz', inc=None, dec=None, pmag=None,
             transposed=False, device=0, **kw):
    """
    Dense sensitivity (Jacobian) matrix of ``field`` with respect to the
    physical property of every cell of ``prisms``.

    Column ``q`` is the field of cell ``q`` alone with unit density (gravity),
    or with magnetization ``pmag`` (magnetic fields; for ``'tf'`` pass the
    inducing direction ``inc, dec``; ``pmag`` defaults to the unit vector along
    it, the harvester convention, harvester.py:958-962).  Masked cells
    (``None``) give zero columns, like the reference's per-cell loop
    (eqlayer.py:108-111, imaging.py:116-117).  Shape ``(npoints, ncells)``, or
    its transpose with ``transposed=True``.
    """
    _check_points(xp, yp, zp, 'shape')
    fvec = None if inc is None else utils.dircos(inc, dec)
    if field in MAGNETIC_FIELDS:
        if field == 'tf' and fvec is None:
            raise ValueError("'tf' needs the inducing field direction inc, dec")
        if pmag is None:
            if fvec is None:
                raise ValueError("magnetic Jacobians need pmag or inc, dec")
            pmag = fvec
        pmag = _pmag_vector(pmag, fvec)
    flat = flatten(prisms, None, override=True)
    with Model(flat, None, device=device) as model:
        jac = model.jacobian(xp, yp, zp, field, pmag=pmag, fvec=fvec,
                             transposed=transposed, **kw)
    if flat.size == flat.total:
        return jac
    # scatter the kept cells' columns back to mesh positions
    if transposed:
        full = numpy.zeros((flat.total, xp.size))
        full[flat.index, :] = jac
    else:
        full = numpy.zeros((xp.size, flat.total))
        full[:, flat.index] = jac
    return full


def adjoint(xp, yp, zp, prisms, data, field='gz', inc=None, dec=None, pmag=None, device=0, **kw):
    """
    Transposed sensitivity applied to a data vector, ``jacobian(...).T @ data``,
    computed on the fly (nothing of size npoints x ncells is stored).

    The reference builds this product layer by layer in ``imaging.migrate``
    (``dot(sensibility_T, gz)``, imaging.py:116-118) and it is the gradient
    product of a linear misfit (misfit.py:280-294).  One value per entry of
    ``prisms``; masked cells (``None``) get 0.  Arguments as :func:`jacobian`.
    """
    _check_points(xp, yp, zp, 'shape')
    fvec = None if inc is None else utils.dircos(inc, dec)
    if field in MAGNETIC_FIELDS:
        if field == 'tf' and fvec is None:
            raise ValueError("'tf' needs the inducing field direction inc, dec")
        if pmag is None:
            if fvec is None:
                raise ValueError("magnetic adjoints need pmag or inc, dec")
            pmag = fvec
        pmag = _pmag_vector(pmag, fvec)
    flat = flatten(prisms, None, override=True)
    with Model(flat, None, device=device) as model:
        got = model.adjoint(xp, yp, zp, data, field, pmag=pmag, fvec=fvec, **kw)
    if flat.size == flat.total:
        return got
    full = numpy.zeros(flat.total)
    full[flat.index] = got
    return full


# --------------------------------------------------------------------------
# the drop-in API: gravity
# --------------------------------------------------------------------------
def potential(xp, yp, zp, prisms, dens=None):
    """
    Calculates the gravitational potential (SI units!).

    * xp, yp, zp : float64 arrays with the coordinates of the computation points
    * prisms : list of ``Prism`` (``None`` entries and prisms without the
      ``'density'`` property are ignored), a ``PrismMesh`` or a ``PrismRelief``
    * dens : float or None; if given, used instead of every prism's density

    Returns the field on xp, yp, zp.  Reference: prism.py:98-143.
    """
    return _gravity(['potential'], xp, yp, zp, prisms, dens)['potential']


def gx(xp, yp, zp, prisms, dens=None):
    """x (North) component of gravity in mGal.  Parameters as :func:`potential`.
    Reference: prism.py:146-191."""
    return _gravity(['gx'], xp, yp, zp, prisms, dens)['gx']


def gy(xp, yp, zp, prisms, dens=None):
    """y (East) component of gravity in mGal.  Parameters as :func:`potential`.
    Reference: prism.py:194-239."""
    return _gravity(['gy'], xp, yp, zp, prisms, dens)['gy']


def gz(xp, yp, zp, prisms, dens=None):
    """z (Down) component of gravity in mGal.  Parameters as :func:`potential`.
    Reference: prism.py:242-287."""
    return _gravity(['gz'], xp, yp, zp, prisms, dens)['gz']


def gxx(xp, yp, zp, prisms, dens=None):
    """xx component of the gravity gradient tensor in Eotvos.  Parameters as
    :func:`potential`.  Reference: prism.py:290-335."""
    return _gravity(['gxx'], xp, yp, zp, prisms, dens)['gxx']


def gxy(xp, yp, zp, prisms, dens=None):
    """xy component of the gravity gradient tensor in Eotvos (with the
    reference's singularity shift).  Reference: prism.py:338-391."""
    return _gravity(['gxy'], xp, yp, zp, prisms, dens)['gxy']


def gxz(xp, yp, zp, prisms, dens=None):
    """xz component of the gravity gradient tensor in Eotvos (with the
    reference's singularity shift).  Reference: prism.py:394-447."""
    return _gravity(['gxz'], xp, yp, zp, prisms, dens)['gxz']


def gyy(xp, yp, zp, prisms, dens=None):
    """yy component of the gravity gradient tensor in Eotvos.
    Reference: prism.py:450-495."""
    return _gravity(['gyy'], xp, yp, zp, prisms, dens)['gyy']


def gyz(xp, yp, zp, prisms, dens=None):
    """yz component of the gravity gradient tensor in Eotvos (with the
    reference's singularity shift).  Reference: prism.py:498-551."""
    return _gravity(['gyz'], xp, yp, zp, prisms, dens)['gyz']


def gzz(xp, yp, zp, prisms, dens=None):
    """zz component of the gravity gradient tensor in Eotvos.
    Reference: prism.py:554-599."""
    return _gravity(['gzz'], xp, yp, zp, prisms, dens)['gzz']


# --------------------------------------------------------------------------
# the drop-in API: magnetic
# --------------------------------------------------------------------------
def tf(xp, yp, zp, prisms, inc, dec, pmag=None):
    """
    Total-field magnetic anomaly in nT.

    * prisms : prisms without the ``'magnetization'`` property are ignored; a
      scalar magnetization is an intensity along the inducing field
    * inc, dec : inclination and declination of the regional field (degrees)
    * pmag : magnetization vector (or scalar intensity) used instead of the
      prisms' own

    Reference: prism.py:602-662.
    """
    fvec = utils.dircos(inc, dec)
    return _magnetic(['tf'], xp, yp, zp, prisms, pmag, fvec, 'length')['tf']


def bx(xp, yp, zp, prisms, pmag=None):
    """x component of the magnetic induction in nT; ``'magnetization'`` must be
    a vector.  Reference: prism.py:665-708."""
    return _magnetic(['bx'], xp, yp, zp, prisms, pmag, None, 'shape')['bx']


def by(xp, yp, zp, prisms, pmag=None):
    """y component of the magnetic induction in nT; ``'magnetization'`` must be
    a vector.  Reference: prism.py:711-754."""
    return _magnetic(['by'], xp, yp, zp, prisms, pmag, None, 'shape')['by']


def bz(xp, yp, zp, prisms, pmag=None):
    """z component of the magnetic induction in nT; ``'magnetization'`` must be
    a vector.  Reference: prism.py:757-800."""
    return _magnetic(['bz'], xp, yp, zp, prisms, pmag, None, 'shape')['bz']


# --------------------------------------------------------------------------
# the drop-in API: second derivatives of the 1/r volume integral
# --------------------------------------------------------------------------
def _kernel(name, xp, yp, zp, prism):
    # the reference calls _prism.g<comp> with density 1 and applies no unit
    # factor (prism.py:828-835 ...), so xy/xz/yz inherit the singularity shift
    return _gravity([name], xp, yp, zp, [prism], 1, 'shape', scaled=False)[name]


def kernelxx(xp, yp, zp, prism):
    """xx derivative of the 1/r volume integral of one prism.
    Reference: prism.py:803-835."""
    return _kernel('gxx', xp, yp, zp, prism)


def kernelyy(xp, yp, zp, prism):
    """yy derivative.  Reference: prism.py:838-870."""
    return _kernel('gyy', xp, yp, zp, prism)


def kernelzz(xp, yp, zp, prism):
    """zz derivative.  Reference: prism.py:873-905."""
    return _kernel('gzz', xp, yp, zp, prism)


def kernelxy(xp, yp, zp, prism):
    """xy derivative.  Reference: prism.py:908-940."""
    return _kernel('gxy', xp, yp, zp, prism)


def kernelxz(xp, yp, zp, prism):
    """xz derivative.  Reference: prism.py:943-975."""
    return _kernel('gxz', xp, yp, zp, prism)


def kernelyz(xp, yp, zp, prism):
    """yz derivative.  Reference: prism.py:978-1010."""
    return _kernel('gyz', xp, yp, zp, prism)
