"""
3D potential-field inversion by planting anomalous densities -- the
``fatiando.gravmag.harvester`` API on the B200 effect engine.

Same public surface as the reference (harvester.py:66-962): :func:`harvest`,
:func:`iharvest`, :func:`sow`, :func:`loadseeds`, :func:`weights`, the data
containers :class:`Potential`, :class:`Gz`, :class:`Gxx` ... :class:`Gzz`,
:class:`TotalField`, and :class:`Seed`-like / :class:`Neighbor` records.  The
growth algorithm (Uieda & Barbosa 2012) and its bookkeeping -- which cells are
candidates, the compactness regulariser, the acceptance rule -- are host logic
and stay in Python.  What the reference spends its time on does not:

* the *effect* of every new candidate cell on every data set
  (``Data.effect``, harvester.py:690-694; one ``prism.<field>`` call per cell
  and data set, :487-493) is evaluated on the GPU in one launch per data set
  and stays in device memory;
* the per-accretion loop over all candidates that forms ``predicted + effect``
  and reduces it to a data misfit and a shape-of-anomaly value
  (``_grow`` :412-431 with ``_misfitfunc`` :447-456 and ``_shapefunc``
  :434-444) is ONE kernel; two doubles per candidate and data set come back.

``predicted`` keeps the reference's float32 storage (harvester.py:401).  Only
prism meshes are supported (the tesseroid engine is outside this package).
There is no CPU path: without the CUDA library or a device :func:`harvest`
raises.
"""
import bisect
import ctypes
import json
from math import sqrt

import numpy

from . import _lib, utils
from . import prism as prism_engine
from .mesher import Prism

_DIRECTIONS = ('above', 'below', 'north', 'south', 'east', 'west')


# --------------------------------------------------------------------------
# seeds
# --------------------------------------------------------------------------
def loadseeds(fname):
    """
    Read seed locations and physical properties from a JSON file (name or open
    file): ``[[x, y, z, {"density": d, ...}], ...]``.  Pass the result to
    :func:`sow`.  (harvester.py:78-142)
    """
    if isinstance(fname, str):
        with open(fname) as handle:
            return json.load(handle)
    return json.load(fname)


class PrismSeed(Prism):
    """A seed: the mesh cell that holds the seed location, with the seed's props."""

    def __init__(self, i, location, prism, props):
        Prism.__init__(self, prism.x1, prism.x2, prism.y1, prism.y2, prism.z1, prism.z2,
                       props=props)
        self.i = i
        self.seed = i
        self.x, self.y, self.z = location


def _find_index(point, mesh):
    """Index of the unmasked cell containing ``point``, or None (harvester.py:192-221)."""
    x1, x2, y1, y2, z1, z2 = mesh.bounds
    nz, ny, nx = mesh.shape
    x, y, z = point
    inside_xy = x1 <= x <= x2 and y1 <= y <= y2
    inside_z = (z1 <= z <= z2) if mesh.zdown else (z2 <= z <= z1)
    if not (inside_xy and inside_z):
        return None
    zs = mesh.get_zs()
    if mesh.zdown:
        k = bisect.bisect_left(zs, z) - 1
    else:
        k = len(zs) - bisect.bisect_left(zs[::-1], z)
    j = bisect.bisect_left(mesh.get_ys(), y) - 1
    i = bisect.bisect_left(mesh.get_xs(), x) - 1
    index = i + j * nx + k * nx * ny
    if mesh[index] is None:
        return None
    return index


def sow(locations, mesh):
    """
    Create seeds from ``[[x, y, z, props], ...]``; locations falling into a cell
    that already holds a seed are dropped.  (harvester.py:145-189)
    """
    if getattr(mesh, 'celltype', None) is not Prism and \
            getattr(getattr(mesh, 'celltype', None), '__name__', '') != 'Prism':
        raise ValueError("only prism meshes are supported by the B200 harvester")
    seeds = []
    taken = set()
    for x, y, z, props in locations:
        index = _find_index((x, y, z), mesh)
        if index is None:
            raise ValueError("Couldn't find seed at location (%g,%g,%g)" % (x, y, z))
        if index not in taken:
            taken.add(index)
            seeds.append(PrismSeed(index, (x, y, z), mesh[index], props))
    return seeds


def weights(x, y, seeds, influences, decay=2):
    """
    Data weights that fade with the distance to the nearest seed
    (harvester.py:619-648): ``exp(-min_s(((x-xs)^2 + (y-ys)^2)/influence_s^2)^decay)``.
    """
    distances = numpy.array([((x - s.x) ** 2 + (y - s.y) ** 2) / influence ** 2
                             for s, influence in zip(seeds, influences)])
    return numpy.exp(-(distances.min(axis=0) ** decay))


# --------------------------------------------------------------------------
# data containers
# --------------------------------------------------------------------------
class Data(object):
    """Observed data of one kind on one set of points (harvester.py:651-668)."""
    field = None
    prop = None

    def __init__(self, x, y, z, data, weights, meshtype):
        if meshtype != 'prism':
            if meshtype == 'tesseroid':
                raise AttributeError("the B200 harvester supports prism meshes only")
            raise AttributeError("Invalid mesh type '%s'" % (meshtype))
        self.x, self.y, self.z = x, y, z
        self.observed = data
        self.size = len(data)
        self.norm = numpy.linalg.norm(data)
        self.meshtype = meshtype
        self.engine = prism_engine
        self.weights = weights

    def effect(self, prism, props):
        """Field of one cell with physical properties ``props`` at the data points."""
        if self.prop not in props:
            return numpy.zeros(self.size, dtype='f')
        return getattr(self.engine, self.field)(self.x, self.y, self.z, [prism], props[self.prop])


class Potential(Data):
    """Gravitational potential data (harvester.py:671-694)."""
    field = 'potential'
    prop = 'density'

    def __init__(self, x, y, z, data, weights=1., meshtype='prism'):
        Data.__init__(self, x, y, z, data, weights, meshtype)
        self.effectfunc = getattr(self.engine, self.field)


class Gz(Potential):
    """Gravity anomaly data."""
    field = 'gz'


class Gxx(Potential):
    """North-north gravity gradient data."""
    field = 'gxx'


class Gxy(Potential):
    """North-east gravity gradient data."""
    field = 'gxy'


class Gxz(Potential):
    """North-vertical gravity gradient data."""
    field = 'gxz'


class Gyy(Potential):
    """East-east gravity gradient data."""
    field = 'gyy'


class Gyz(Potential):
    """East-vertical gravity gradient data."""
    field = 'gyz'


class Gzz(Potential):
    """Vertical-vertical gravity gradient data."""
    field = 'gzz'


class TotalField(Potential):
    """Total-field magnetic anomaly data (harvester.py:930-962)."""
    field = 'tf'
    prop = 'magnetization'

    def __init__(self, x, y, z, data, inc, dec, weights=1., meshtype='prism'):
        if meshtype != 'prism':
            raise AttributeError(
                "Unsupported mesh type '%s' for total field anomaly." % (meshtype))
        Potential.__init__(self, x, y, z, data, weights, meshtype)
        self.inc = inc
        self.dec = dec

    def effect(self, prism, props):
        if self.prop not in props:
            return numpy.zeros(self.size, dtype='f')
        return self.engine.tf(self.x, self.y, self.z, [prism], self.inc, self.dec,
                              pmag=props[self.prop])


# --------------------------------------------------------------------------
# the device engine
# --------------------------------------------------------------------------
class EffectEngine(object):
    """
    Device-resident state of one inversion: the data sets, ``predicted`` and one
    effect row per live candidate (csrc/harvest.cu).
    """

    _handle = None

    def __init__(self, data, device=0, reference_order=False):
        _lib.require_device()
        self.sizes = [int(d.size) for d in data]
        self.offsets = numpy.concatenate([[0], numpy.cumsum(self.sizes)]).astype(numpy.int64)
        self.ndata = len(data)
        self.ntot = int(self.offsets[-1])
        self.props = [d.prop for d in data]
        cat = lambda parts: numpy.ascontiguousarray(numpy.concatenate(
            [numpy.asarray(p, dtype=numpy.float64).ravel() for p in parts]))
        x, y, z = cat([d.x for d in data]), cat([d.y for d in data]), cat([d.z for d in data])
        obs = cat([d.observed for d in data])
        w = cat([numpy.broadcast_to(numpy.asarray(d.weights, dtype=numpy.float64), (d.size,))
                 for d in data])
        fields = (ctypes.c_int * self.ndata)(*[_lib.FIELD_ID[d.field] for d in data])
        sizes = (ctypes.c_int64 * self.ndata)(*self.sizes)
        fvec = numpy.zeros((self.ndata, 3))
        for k, d in enumerate(data):
            if d.field in prism_engine.MAGNETIC_FIELDS:
                fvec[k] = utils.dircos(d.inc, d.dec)
        scale = numpy.array([prism_engine.unit_scale(d.field) for d in data])
        self._flags = _lib.REFERENCE_ORDER if reference_order else 0
        self._handle = ctypes.c_void_p()
        _lib.check(_lib.lib().fb200_harvest_create(
            device, self.ndata, sizes, fields, _lib.dptr(x), _lib.dptr(y), _lib.dptr(z),
            _lib.dptr(obs), _lib.dptr(w), _lib.dptr(fvec), _lib.dptr(scale),
            ctypes.byref(self._handle)))
        self._free = []
        self._next = 0

    def close(self):
        if self._handle:
            _lib.lib().fb200_harvest_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    __del__ = close

    # -- effect rows --------------------------------------------------------
    def _take_row(self):
        if self._free:
            return self._free.pop()
        self._next += 1
        return self._next - 1

    def release(self, row):
        if self._handle:
            self._free.append(row)

    def effects(self, cells, props):
        """Evaluate the effect rows of ``cells`` (all with ``props``); returns their rows."""
        n = len(cells)
        rows = numpy.array([self._take_row() for _ in range(n)], dtype=numpy.int64)
        if n == 0:
            return rows
        bounds = numpy.array([[c.x1, c.x2, c.y1, c.y2, c.z1, c.z2] for c in cells],
                             dtype=numpy.float64)
        dens = numpy.full(n, float(props['density']) if 'density' in props else numpy.nan)
        mag = numpy.zeros((n, 3))
        kind = 0
        if 'magnetization' in props:
            value = props['magnetization']
            if isinstance(value, (float, int)):
                kind, mag[:, 0] = 1, value
            else:
                kind, mag[:] = 2, numpy.asarray(value, dtype=numpy.float64)
        kinds = (ctypes.c_int * n)(*([kind] * n))
        _lib.check(_lib.lib().fb200_harvest_effects(
            self._handle, n, _lib.dptr(bounds), _lib.dptr(dens), _lib.dptr(mag), kinds,
            rows.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), self._flags))
        return rows

    def accrete(self, row):
        _lib.check(_lib.lib().fb200_harvest_accrete(self._handle, int(row)))

    def evaluate(self, rows):
        """(misfit, shape) arrays of shape (len(rows), ndata); row -1 = no effect."""
        rows = numpy.ascontiguousarray(rows, dtype=numpy.int64)
        misfit = numpy.empty((rows.size, self.ndata))
        shape = numpy.empty((rows.size, self.ndata))
        # at most 65535 candidates per launch
        for a in range(0, rows.size, 65535):
            part = rows[a:a + 65535]
            _lib.check(_lib.lib().fb200_harvest_evaluate(
                self._handle, part.size, part.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                _lib.dptr(misfit[a:a + 65535]), _lib.dptr(shape[a:a + 65535])))
        return misfit, shape

    def _split(self, flat):
        return [flat[a:b].copy() for a, b in zip(self.offsets[:-1], self.offsets[1:])]

    def predicted(self):
        out = numpy.empty(self.ntot, dtype=numpy.float32)
        _lib.check(_lib.lib().fb200_harvest_predicted(
            self._handle, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float))))
        return self._split(out)

    def effect(self, row):
        out = numpy.empty(self.ntot)
        _lib.check(_lib.lib().fb200_harvest_effect(self._handle, int(row), _lib.dptr(out)))
        parts = self._split(out)
        return parts

    def launches(self):
        return int(_lib.lib().fb200_harvest_launches(self._handle))


class Neighbor(object):
    """
    A candidate cell: mesh index ``i``, the ``props`` it would get, the ``seed``
    it belongs to, its ``distance`` to that seed in cells, and its ``effect``
    (list of arrays, one per data set -- fetched from the device on access).
    """

    def __init__(self, i, props, seed, distance, engine, row):
        self.i = i
        self.props = props
        self.seed = seed
        self.distance = distance
        self._engine = engine
        self._row = row

    @property
    def effect(self):
        parts = self._engine.effect(self._row)
        # data sets whose property the candidate lacks: float32 zeros, like the reference
        return [p if prop in self.props else numpy.zeros(p.size, dtype='f')
                for p, prop in zip(parts, self._engine.props)]

    def __del__(self):
        try:
            self._engine.release(self._row)
        except Exception:
            pass


# --------------------------------------------------------------------------
# neighbourhood bookkeeping (host logic, harvester.py:459-581)
# --------------------------------------------------------------------------
def _test_restriction(restrict):
    if restrict is None:
        return []
    for case in restrict:
        if case not in _DIRECTIONS:
            raise ValueError("Unrecognized item in restrict: %s" % case)
    return restrict


def _index2ijk(index, mesh):
    nz, ny, nx = mesh.shape
    k, rest = divmod(index, nx * ny)
    j, i = divmod(rest, nx)
    return i, j, k


def _distance(n, m, mesh):
    ni, nj, nk = _index2ijk(n, mesh)
    mi, mj, mk = _index2ijk(m, mesh)
    return sqrt((ni - mi) ** 2 + (nj - mj) ** 2 + (nk - mk) ** 2)


def _masked_cells(mesh):
    """The mesh's masked cells as a set (``mesh[i] is None`` without building a Prism);
    cached on the mesh and refreshed when its mask list changes."""
    mask = mesh.mask
    cache = getattr(mesh, '_fb200_masked', None)
    if cache is None or cache[0] is not mask or cache[1] != len(mask):
        cache = (mask, len(mask), frozenset(mask))
        mesh._fb200_masked = cache
    return cache[2]


def _neighbor_indexes(n, mesh, restrict):
    """
    The up-to-six face neighbours of cell ``n`` that exist and are not masked,
    in the reference's order: above, below, north, south, east, west
    (harvester.py:539-581, including its ``> 0`` test for the cell above).
    """
    nz, ny, nx = mesh.shape
    layer = nx * ny
    inlayer = n % layer
    candidates = (
        ('above', n - layer, n - layer > 0),
        ('below', n + layer, n + layer < mesh.size),
        ('north', n + 1, n % nx < nx - 1),
        ('south', n - 1, n % nx != 0),
        ('east', n + nx, inlayer < nx * (ny - 1)),
        ('west', n - nx, inlayer >= nx),
    )
    masked = _masked_cells(mesh)
    return [idx for name, idx, ok in candidates
            if name not in restrict and ok and idx not in masked]


class _CellBounds(object):
    """The six bounds of an unmasked mesh cell, by the expressions of ``PrismMesh.__getitem__``
    (mesh.py:629-634) without the ``Prism`` object and its props."""
    __slots__ = ('x1', 'x2', 'y1', 'y2', 'z1', 'z2')

    def __init__(self, mesh, index):
        nz, ny, nx = mesh.shape
        k, rest = divmod(index, nx * ny)
        j, i = divmod(rest, nx)
        dx, dy, dz = mesh.dims
        self.x1 = mesh.bounds[0] + dx * i
        self.y1 = mesh.bounds[2] + dy * j
        self.z1 = mesh.bounds[4] + dz * k
        self.x2, self.y2, self.z2 = self.x1 + dx, self.y1 + dy, self.z1 + dz


def _shares_prop(props, other):
    return any(p in other for p in props)


def _is_neighbor(index, props, neighborhood):
    return any(index in nbrs and _shares_prop(props, nbrs[index].props) for nbrs in neighborhood)


def _in_estimate(index, props, estimate):
    return index in estimate and _shares_prop(props, estimate[index])


class _Candidates(dict):
    """
    ``{cell index: Neighbor}`` of one seed, in insertion order like the reference's dicts,
    that also keeps the effect rows and the distances of its values as numpy arrays in the
    same order: ``_grow`` hands them to the device as they are instead of walking thousands of
    Python objects per accretion.  ``pop`` / ``update`` / item assignment and deletion keep the
    arrays in step.
    """

    def __init__(self, items=()):
        dict.__init__(self)
        self.cells = numpy.empty(0, dtype=numpy.int64)
        self.rows = numpy.empty(0, dtype=numpy.int64)
        self.distances = numpy.empty(0, dtype=numpy.float64)
        self.update(items)

    def __setitem__(self, key, value):
        if key in self:
            at = int(numpy.flatnonzero(self.cells == key)[0])
            self.rows[at], self.distances[at] = value._row, value.distance
        else:
            self.cells = numpy.append(self.cells, key)
            self.rows = numpy.append(self.rows, value._row)
            self.distances = numpy.append(self.distances, value.distance)
        dict.__setitem__(self, key, value)

    def update(self, other=(), **kw):
        items = list(dict(other, **kw).items())
        fresh = [(k, v) for k, v in items if k not in self]
        for k, v in items:
            if k in self:
                self[k] = v
        if fresh:
            self.cells = numpy.concatenate([self.cells, numpy.array([k for k, _ in fresh], dtype=numpy.int64)])
            self.rows = numpy.concatenate([self.rows, numpy.array([v._row for _, v in fresh], dtype=numpy.int64)])
            self.distances = numpy.concatenate([self.distances,
                                                numpy.array([v.distance for _, v in fresh], dtype=numpy.float64)])
            for k, v in fresh:
                dict.__setitem__(self, k, v)

    def __delitem__(self, key):
        dict.__delitem__(self, key)
        keep = self.cells != key
        self.cells, self.rows, self.distances = self.cells[keep], self.rows[keep], self.distances[keep]

    def pop(self, key, *default):
        if key not in self:
            return dict.pop(self, key, *default)
        value = dict.__getitem__(self, key)
        del self[key]
        return value

    def popitem(self):
        key = next(reversed(self))
        return key, self.pop(key)

    def clear(self):
        dict.clear(self)
        self.cells, self.rows, self.distances = self.cells[:0], self.rows[:0], self.distances[:0]

    def setdefault(self, key, default=None):
        if key not in self:
            self[key] = default
        return dict.__getitem__(self, key)


def _get_neighbors(cell, neighborhood, estimate, mesh, engine, restrict):
    """New candidates around ``cell`` (a seed or an accreted Neighbor), with effect rows."""
    indexes = [n for n in _neighbor_indexes(cell.i, mesh, restrict)
               if not _is_neighbor(n, cell.props, neighborhood)
               and not _in_estimate(n, cell.props, estimate)]
    rows = engine.effects([_CellBounds(mesh, i) for i in indexes], cell.props)
    return _Candidates((i, Neighbor(i, cell.props, cell.seed, _distance(i, cell.seed, mesh), engine, int(r)))
                       for i, r in zip(indexes, rows))


def _totals(values):
    """Sum over the data sets in list order, like the reference's ``result +=`` loops."""
    total = numpy.zeros(values.shape[0])
    for d in range(values.shape[1]):
        total += values[:, d]
    return total


def _grow(neighbors, engine, totalmisfit, mu, regularizer, threshold):
    """
    The candidate that lowers the misfit by at least ``threshold`` and has the
    smallest goal function; first one wins ties (harvester.py:412-431).
    """
    if not neighbors:
        return None, None, None, None
    if isinstance(neighbors, _Candidates):
        rows, distances, cells = neighbors.rows, neighbors.distances, neighbors.cells
    else:                                   # a plain dict handed in by a caller
        cands = list(neighbors.values())
        rows = numpy.fromiter((c._row for c in cands), dtype=numpy.int64, count=len(cands))
        distances = numpy.fromiter((c.distance for c in cands), dtype=numpy.float64, count=len(cands))
        cells = numpy.fromiter(neighbors.keys(), dtype=numpy.int64, count=len(cands))
    misfit, shape = engine.evaluate(rows)
    misfit, shape = _totals(misfit), _totals(shape)
    ok = (misfit < totalmisfit) & (numpy.abs(misfit - totalmisfit) / totalmisfit >= threshold)
    if not ok.any():
        return None, None, None, None
    reg = regularizer + distances
    goal = numpy.where(ok, shape + mu * reg, numpy.inf)
    best = int(numpy.argmin(goal))          # first occurrence of the minimum
    return neighbors[int(cells[best])], float(goal[best]), float(misfit[best]), float(reg[best])


def fmt_estimate(estimate, size):
    """``{prop: SparseList}`` from the per-cell props dicts (harvester.py:395-409)."""
    output = {}
    for i in estimate:
        for p, value in estimate[i].items():
            if p not in output:
                output[p] = utils.SparseList(size)
            output[p][i] = value
    return output


def _run(data, seeds, mesh, compactness, threshold, restrict, device, reference_order,
         fetch_predicted, final=None, engine_factory=None):
    for d in data:
        if not isinstance(d, Data):
            raise TypeError("data must be harvester data containers (Gz, Gzz, TotalField, ...)")
    # engine_factory exists so that the bookkeeping above can be tested on machines
    # without a GPU (tests inject a numpy engine); the product default is the CUDA engine
    if engine_factory is None:
        engine = EffectEngine(data, device=device, reference_order=reference_order)
    else:
        engine = engine_factory(data)
    try:
        nseeds = len(seeds)
        estimate = dict((s.i, s.props) for s in seeds)
        neighbors = []
        for seed in seeds:
            neighbors.append(_get_neighbors(seed, neighbors, estimate, mesh, engine, restrict))
        # effect of the seeds themselves (harvester.py:397-405)
        for seed in seeds:
            row = engine.effects([mesh[seed.i]], seed.props)[0]
            engine.accrete(row)
            engine.release(int(row))
        misfit, shape = engine.evaluate([-1])
        totalgoal = float(_totals(shape)[0])
        totalmisfit = float(_totals(misfit)[0])
        regularizer = 0.
        mu = compactness * 1 / (sum(mesh.shape) / 3)
        predicted = engine.predicted() if fetch_predicted else None
        yield [estimate, predicted, None, neighbors, totalgoal, totalmisfit, regularizer, engine]
        for iteration in range(mesh.size - nseeds):
            grew = False
            for s in range(nseeds):
                best, bestgoal, bestmisfit, bestreg = _grow(
                    neighbors[s], engine, totalmisfit, mu, regularizer, threshold)
                if best is None:
                    continue
                estimate.setdefault(best.i, {}).update(best.props)
                totalgoal, totalmisfit, regularizer = bestgoal, bestmisfit, bestreg
                engine.accrete(best._row)
                neighbors[s].pop(best.i)
                neighbors[s].update(_get_neighbors(best, neighbors, estimate, mesh, engine, restrict))
                grew = True
                if fetch_predicted:
                    for p, new in zip(predicted, engine.predicted()):
                        p[:] = new
                yield [estimate, predicted, best, neighbors, totalgoal, totalmisfit, regularizer,
                       engine]
                del best
            if not grew:
                break
        if final is not None:
            final['predicted'] = engine.predicted()
            final['gpu_launches'] = engine.launches()
    finally:
        # the Neighbor records may outlive the engine; their rows just lapse
        engine.close()


def iharvest(data, seeds, mesh, compactness, threshold, restrict, device=0,
             reference_order=False, engine_factory=None):
    """
    Step through the inversion one accretion at a time.  Yields
    ``[estimate, predicted, new, neighbors, goal, misfit, regularizer]``; the
    first yield holds the seeds only (``new`` is None).  (harvester.py:331-392)
    """
    for update in _run(data, seeds, mesh, compactness, threshold, restrict, device,
                       reference_order, True, engine_factory=engine_factory):
        yield update[:7]


def harvest(data, seeds, mesh, compactness, threshold, report=False, restrict=None, device=0,
            reference_order=False, engine_factory=None):
    """
    Run the planting inversion (harvester.py:224-328).

    * data : list of data containers (:class:`Gz`, :class:`Gzz`, ...)
    * seeds : list of seeds from :func:`sow`
    * mesh : :class:`fatiando_b200.mesher.PrismMesh`
    * compactness : weight of the compactness regulariser (> 0)
    * threshold : smallest relative misfit reduction an accretion must bring
    * report : also return ``{'goal', 'shape-of-anomaly', 'misfit', 'regularizer',
      'accretions'}``
    * restrict : directions in which seeds may not grow, any of ``'above'``,
      ``'below'``, ``'north'``, ``'south'``, ``'east'``, ``'west'``

    Returns ``[estimate, predicted]`` (+ the report): ``estimate`` maps each
    physical property to a sparse per-cell list, ``predicted`` is the list of
    predicted data vectors (float32) in the order of ``data``.
    """
    restrict = _test_restriction(restrict)
    accretions = -1
    update, final = None, {}
    for update in _run(data, seeds, mesh, compactness, threshold, restrict, device,
                       reference_order, False, final, engine_factory):
        accretions += 1
    estimate, goal, misfit, regul = update[0], update[4], update[5], update[6]
    output = [fmt_estimate(estimate, mesh.size), final['predicted']]
    if report:
        soa = goal - compactness * 1 / (sum(mesh.shape) / 3) * regul
        output.append({'goal': goal, 'misfit': misfit, 'regularizer': regul,
                       'accretions': accretions, 'shape-of-anomaly': soa,
                       'gpu_launches': final['gpu_launches']})
    return output
