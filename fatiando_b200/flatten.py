"""
The mesh flattener: any "list of prisms" -> SoA arrays ready for upload.

The reference walks its models one ``Prism`` object at a time
(prism.py:131-141; ``PrismMesh.__getitem__`` mesh.py:617-636,
``PrismRelief.__getitem__`` mesh.py:442-458), about 3-7 microseconds of Python
per cell and per field call.  Here a whole model becomes six bound arrays plus
property arrays in one vectorised pass.  The coordinate expressions are the
reference's own, evaluated element-wise in float64, so every bound is the very
same double the reference would hand to its kernel:

* PrismMesh:   ``x1 = bounds[0] + dx*i``, ``x2 = x1 + dx``  (mesh.py:629-634)
* PrismRelief: ``x1 = xc - 0.5*dx`` ..., z ordered against ``ref`` (:447-456)

Skipping rules (prism.py:132, :647-649): ``None`` entries (masked cells) are
always dropped; cells lacking the required property are dropped unless an
override (``dens=`` / ``pmag=``) is given, in which case every non-``None`` cell
counts.  Order is preserved -- it is the summation order.
"""
import numpy as np

BOUND_NAMES = ('x1', 'x2', 'y1', 'y2', 'z1', 'z2')


class FlatModel(object):
    """
    SoA view of a prism model.

    * bounds : list of six float64 arrays (x1, x2, y1, y2, z1, z2), length m
    * density : float64 array (m,) or None
    * magnetization : float64 array (m, 3) or None
    * index : int64 array (m,), position of each kept prism in the input
    * total : number of entries of the input (kept + dropped)
    * nodes : :class:`MeshNodes` when the input was a regular ``PrismMesh``
      carrying the property (its node-sharing form), else ``None``
    * surface : :class:`ReliefSurface` when the input was a ``PrismRelief`` on a
      tight regular grid carrying a density, else ``None``
    """

    def __init__(self, bounds, density, magnetization, index, total, nodes=None, surface=None):
        self.bounds = [np.ascontiguousarray(b, dtype=np.float64) for b in bounds]
        self.density = None if density is None else np.ascontiguousarray(density, dtype=np.float64)
        self.magnetization = (None if magnetization is None
                              else np.ascontiguousarray(magnetization, dtype=np.float64))
        self.index = np.ascontiguousarray(index, dtype=np.int64)
        self.total = int(total)
        self.nodes = nodes
        self.surface = surface

    @property
    def size(self):
        return int(self.bounds[0].size)

    def extent(self):
        """Largest absolute coordinate and smallest non-zero edge (for the
        fast-path range guard)."""
        if self.size == 0:
            return 0.0, np.inf
        hint = getattr(self, 'extent_hint', None)
        if hint is not None:            # set by the flattener where the geometry says it all
            return hint
        big = 0.0
        for b in self.bounds:           # min / max reductions: no temporaries
            big = max(big, -float(b.min()), float(b.max()))
        small = np.inf
        for lo, hi in ((0, 1), (2, 3), (4, 5)):
            e = self.bounds[hi] - self.bounds[lo]
            np.abs(e, out=e)
            e = e[e > 0]
            if e.size:
                small = min(small, float(e.min()))
        return big, small


def _is_scalar_property(value):
    # the reference's test (prism.py:651): a Python float or int
    # (numpy.float64 is a float subclass and counts)
    return isinstance(value, (float, int))


def _is_prism_mesh(obj):
    return (all(hasattr(obj, a) for a in ('bounds', 'shape', 'dims', 'mask', 'props', 'size'))
            and getattr(getattr(obj, 'celltype', None), '__name__', '') == 'Prism')


def _is_prism_relief(obj):
    return all(hasattr(obj, a) for a in ('ref', 'dx', 'dy', 'x', 'y', 'z', 'props', 'size'))


def _mesh_bounds(mesh):
    nz, ny, nx = mesh.shape
    dx, dy, dz = (float(d) for d in mesh.dims)
    b = mesh.bounds
    i = np.arange(nx, dtype=np.float64)
    j = np.arange(ny, dtype=np.float64)
    k = np.arange(nz, dtype=np.float64)
    x1 = float(b[0]) + dx * i
    y1 = float(b[2]) + dy * j
    z1 = float(b[4]) + dz * k
    x2, y2, z2 = x1 + dx, y1 + dy, z1 + dz
    along_x = lambda a: np.tile(a, ny * nz)
    along_y = lambda a: np.tile(np.repeat(a, nx), nz)
    along_z = lambda a: np.repeat(a, nx * ny)
    return [along_x(x1), along_x(x2), along_y(y1), along_y(y2), along_z(z1), along_z(z2)]


def _relief_bounds(relief):
    x = np.asarray(relief.x, dtype=np.float64)
    y = np.asarray(relief.y, dtype=np.float64)
    z = np.asarray(relief.z, dtype=np.float64)
    dx, dy, ref = float(relief.dx), float(relief.dy), float(relief.ref)
    above = z <= ref
    return [x - 0.5 * dx, x + 0.5 * dx, y - 0.5 * dy, y + 0.5 * dy,
            np.where(above, z, ref), np.where(above, ref, z)]


def _property_arrays(props, size):
    """(density or None, magnetization (m,3)/(m,) or None, needs_generic)"""
    dens = mag = None
    if 'density' in props:
        dens = np.asarray(props['density'])
        if dens.dtype == object or dens.shape != (size,):
            return None, None, True
        dens = np.asarray(dens, dtype=np.float64)
    if 'magnetization' in props:
        mag = np.asarray(props['magnetization'])
        if mag.dtype == object or mag.shape not in ((size,), (size, 3)):
            return None, None, True
        mag = np.asarray(mag, dtype=np.float64)
    return dens, mag, False


class MeshNodes(object):
    """
    Node-sharing form of a regular ``PrismMesh`` (see csrc/prism_mesh.cuh).

    * xn, yn, zn : node coordinates (nx+1, ny+1, nz+1), the reference's own bound
      expressions: ``b0 + d*i`` for i < n and ``(b0 + d*(n-1)) + d`` for the last
      (mesh.py:629-634)
    * cell_size : (dx, dy, dz)
    * weights : list of 1 (density) or 3 (mx, my, mz) float64 arrays of
      ``(nx+1)*(ny+1)*(nz+1)`` node weights, z fastest.  The weight of a node is
      the signed sum of the property of the (up to) eight cells it is a corner
      of: ``+`` where it is the cell's (x2, y2, z2)-type corner, alternating with
      every bound that is the lower one instead ((-1)**(i+j+k), _prism.pyx:104).
      Masked cells count as zero.
    """

    def __init__(self, shape, xn, yn, zn, cell_size, weights, hazard=None):
        self.shape = shape
        self.xn, self.yn, self.zn = xn, yn, zn
        self.cell_size = np.array(cell_size, dtype=np.float64)
        self.weights = weights
        self.hazard = hazard if hazard is not None else [np.zeros(0)] * 3


def _mismatched_walls(upper_of_lower_cell, lower_of_upper_cell):
    """
    Coordinates (both variants) of the interior walls that the reference computes
    differently from their two sides -- see fb200_model_set_hazard_planes.
    """
    a = np.asarray(upper_of_lower_cell, dtype=np.float64)
    b = np.asarray(lower_of_upper_cell, dtype=np.float64)
    bad = a != b
    return np.ascontiguousarray(np.concatenate([a[bad], b[bad]]))


def _node_coordinates(b0, d, n):
    i = np.arange(n, dtype=np.float64)
    lower = float(b0) + float(d) * i
    return np.concatenate([lower, [lower[-1] + float(d)]])


def _node_weights(values, shape):
    """2x2x2 signed-sum stencil of a per-cell property, returned z fastest."""
    nz, ny, nx = shape
    pad = np.zeros((nz + 2, ny + 2, nx + 2))
    pad[1:-1, 1:-1, 1:-1] = values.reshape(nz, ny, nx)
    # sum over (dc, db, da) in {0,1}^3 of (-1)^(da+db+dc) pad[c+dc, b+db, a+da]: along every
    # axis that is (lower neighbour) - (upper neighbour), i.e. minus a forward difference
    w = np.diff(np.diff(np.diff(pad, axis=2), axis=1), axis=0)
    np.negative(w, out=w)
    return np.ascontiguousarray(np.transpose(w, (1, 2, 0))).ravel()     # [b][a][c]


def mesh_geometry(mesh):
    """
    Geometry-only :class:`MeshNodes` (no weights) of a regular ``PrismMesh`` without
    masked cells -- what the node-sharing Jacobian needs -- or ``None``.
    """
    if not _is_prism_mesh(mesh):
        return None
    mask = getattr(mesh, 'mask', None)
    if mask is not None and len(mask):
        return None
    nz, ny, nx = (int(v) for v in mesh.shape)
    dx, dy, dz = (float(d) for d in mesh.dims)
    if not (dx > 0 and dy > 0 and dz > 0):
        return None
    b = mesh.bounds
    nodes = [_node_coordinates(b[0], dx, nx), _node_coordinates(b[2], dy, ny),
             _node_coordinates(b[4], dz, nz)]
    hazard = [_mismatched_walls(c[:-2] + d, c[1:-1]) for c, d in zip(nodes, (dx, dy, dz))]
    return MeshNodes((nz, ny, nx), nodes[0], nodes[1], nodes[2], (dx, dy, dz), [], hazard)


def mesh_nodes(mesh, prop, fvec=None):
    """
    :class:`MeshNodes` of a regular ``PrismMesh`` carrying ``prop`` on every
    cell, or ``None`` when ``mesh`` is not such a mesh (then only the per-cell
    evaluation applies).
    """
    if not _is_prism_mesh(mesh) or prop not in getattr(mesh, 'props', {}):
        return None
    nz, ny, nx = (int(v) for v in mesh.shape)
    size = nx * ny * nz
    dens, mag, generic = _property_arrays(mesh.props, size)
    if generic:
        return None
    dx, dy, dz = (float(d) for d in mesh.dims)
    if not (dx > 0 and dy > 0 and dz > 0):
        return None
    if prop == 'density':
        channels = [dens]
    else:
        if mag.ndim == 1:
            mag = _expand_scalar_magnetization(mag, fvec)
        channels = [mag[:, c] for c in range(3)]
    mask = getattr(mesh, 'mask', None)
    if mask is not None and len(mask):
        keep = np.ones(size)
        keep[np.asarray(mask, dtype=np.int64)] = 0.0
        channels = [ch * keep for ch in channels]
    if not all(np.all(np.isfinite(ch)) for ch in channels):
        return None
    b = mesh.bounds
    nodes = [_node_coordinates(b[0], dx, nx), _node_coordinates(b[2], dy, ny),
             _node_coordinates(b[4], dz, nz)]
    # x2 = x1 + dx of cell i against x1 = b0 + dx*(i+1) of cell i+1 (mesh.py:629-634)
    hazard = [_mismatched_walls(c[:-2] + d, c[1:-1]) for c, d in zip(nodes, (dx, dy, dz))]
    return MeshNodes((nz, ny, nx), nodes[0], nodes[1], nodes[2], (dx, dy, dz),
                     [_node_weights(ch, (nz, ny, nx)) for ch in channels], hazard)


class ReliefSurface(object):
    """
    Surface form of a ``PrismRelief`` on a tight regular grid (csrc/prism_face.cuh).

    * faces : seven float64 arrays ``x1, x2, y1, y2, z_face, z_other, w`` -- one top
      face per cell, ``w = -rho`` with ``rho`` the cell's density once the sign rule
      of ``PrismRelief.addprop`` (mesh.py:484-488) is undone
    * nodes : four float64 arrays ``xn, yn, zn, w`` -- reference-plane corners merged
      over the cells that share them, ``w`` = signed 2x2 sum of ``rho``; zero-weight
      nodes are dropped
    * cell_size : (dx, dy)
    """

    def __init__(self, faces, nodes, cell_size, hazard=None):
        self.faces = [np.ascontiguousarray(a, dtype=np.float64) for a in faces]
        self.nodes = [np.ascontiguousarray(a, dtype=np.float64) for a in nodes]
        self.cell_size = np.array(cell_size, dtype=np.float64)
        self.hazard = hazard if hazard is not None else [np.zeros(0)] * 3


def _grid_axes(x, y):
    """
    If the points (x, y) are the nodes of a rectangular grid, each exactly once, return
    ``(xs, ix, ys, iy)``: the sorted axis coordinates and every point's indices; else None.
    The raveled-meshgrid orders ``gridder.regular`` produces (x slowest, or y slowest) are
    recognised in O(n); anything else goes through a sort.
    """
    n = x.size
    for slow, fast in ((x, y), (y, x)):
        change = np.flatnonzero(slow[1:] != slow[:-1])
        nfast = int(change[0]) + 1 if change.size else n
        if n % nfast:
            continue
        nslow = n // nfast
        s2, f2 = slow.reshape(nslow, nfast), fast.reshape(nslow, nfast)
        sa, fa = s2[:, 0], f2[0]
        if not (np.all(np.diff(sa) > 0) and np.all(np.diff(fa) > 0)):
            continue
        if np.array_equal(s2, np.broadcast_to(sa[:, None], s2.shape)) and \
                np.array_equal(f2, np.broadcast_to(fa[None, :], f2.shape)):
            islow = np.repeat(np.arange(nslow), nfast)
            ifast = np.tile(np.arange(nfast), nslow)
            return (sa, islow, fa, ifast) if slow is x else (fa, ifast, sa, islow)
    xs, ix = np.unique(x, return_inverse=True)
    ys, iy = np.unique(y, return_inverse=True)
    if xs.size * ys.size != n:
        return None
    if np.any(np.bincount(iy * xs.size + ix, minlength=n) != 1):
        return None
    return xs, ix, ys, iy


def relief_surface(relief, prop='density', bounds=None):
    """
    :class:`ReliefSurface` of ``relief``, or ``None`` when it does not apply: not a
    ``PrismRelief`` with a per-cell density, the nodes are not a regular grid, or the
    cells do not tile it (spacing differs from the cell size by more than rounding).
    """
    if prop != 'density' or not _is_prism_relief(relief) or 'density' not in relief.props:
        return None
    n = int(relief.size)
    dens, _, generic = _property_arrays({'density': relief.props['density']}, n)
    if generic or n == 0 or not np.all(np.isfinite(dens)):
        return None
    x = np.asarray(relief.x, dtype=np.float64)
    y = np.asarray(relief.y, dtype=np.float64)
    z = np.asarray(relief.z, dtype=np.float64)
    dx, dy, ref = float(relief.dx), float(relief.dy), float(relief.ref)
    if not (dx > 0 and dy > 0 and np.all(np.isfinite(x)) and np.all(np.isfinite(y))
            and np.all(np.isfinite(z))):
        return None
    grid = _grid_axes(x, y)
    if grid is None:
        return None
    xs, ix, ys, iy = grid
    nx, ny = xs.size, ys.size
    if bounds is None:
        bounds = _relief_bounds(relief)
    # neighbouring cells must share their edge up to rounding: xs[a] + dx/2 == xs[a+1] - dx/2
    tol = 8 * np.finfo(float).eps
    for c, h, big in ((xs, 0.5 * dx, np.abs(xs).max()), (ys, 0.5 * dy, np.abs(ys).max())):
        if c.size > 1 and np.any(np.abs((c[:-1] + h) - (c[1:] - h)) > tol * max(big, h)):
            return None
    rho = np.where(z <= ref, dens, -dens)
    grid = np.zeros((ny + 2, nx + 2))
    grid[iy + 1, ix + 1] = rho
    # node (a, b) is the (x2, y2) corner of cell (a-1, b-1): + ; of (a, b-1) and (a-1, b): - ; (a, b): +
    w = grid[:-1, :-1] - grid[:-1, 1:] - grid[1:, :-1] + grid[1:, 1:]
    xn = np.concatenate([xs - 0.5 * dx, [xs[-1] + 0.5 * dx]])
    yn = np.concatenate([ys - 0.5 * dy, [ys[-1] + 0.5 * dy]])
    b, a = np.nonzero(w)
    nodes = [xn[a], yn[b], np.full(a.size, ref), w[b, a]]
    faces = [bounds[0], bounds[1], bounds[2], bounds[3], z, np.full(n, ref), -rho]
    # xc + dx/2 of one node against xc - dx/2 of the next (mesh.py:447-450)
    hazard = [_mismatched_walls(xs[:-1] + 0.5 * dx, xs[1:] - 0.5 * dx),
              _mismatched_walls(ys[:-1] + 0.5 * dy, ys[1:] - 0.5 * dy), np.zeros(0)]
    return ReliefSurface(faces, nodes, (dx, dy), hazard)


def flatten(prisms, prop=None, override=False, fvec=None):
    """
    Flatten ``prisms`` for a field that needs physical property ``prop``.

    * prop : ``'density'``, ``'magnetization'`` or ``None`` (bounds only)
    * override : a ``dens=`` / ``pmag=`` value replaces the property, so cells
      without it are kept (prism.py:132-137)
    * fvec : direction cosines of the inducing field; scalar magnetizations are
      expanded to ``mag * fvec`` like prism.py:650-653 (total-field anomaly
      only -- without ``fvec`` a scalar magnetization is an error, as it is
      for ``bx, by, bz`` in the reference, prism.py:702)
    """
    bounds = dens = mag = keep = None
    total = None
    if _is_prism_mesh(prisms) or _is_prism_relief(prisms):
        total = int(prisms.size)
        dens, mag, generic = _property_arrays(prisms.props, total)
        if not generic:
            bounds = _mesh_bounds(prisms) if _is_prism_mesh(prisms) else _relief_bounds(prisms)
            keep = np.ones(total, dtype=bool)
            mask = getattr(prisms, 'mask', None)
            if mask is not None and len(mask):
                keep[np.asarray(mask, dtype=np.int64)] = False
            if prop is not None and not override and prop not in prisms.props:
                keep[:] = False
    if bounds is None:
        return _flatten_generic(prisms, prop, override, fvec)
    index = np.nonzero(keep)[0]
    if index.size != total:
        bounds = [b[index] for b in bounds]
        dens = None if dens is None else dens[index]
        mag = None if mag is None else mag[index]
    if override:
        dens = mag = None
    if prop != 'density':
        dens = None
    if prop != 'magnetization':
        mag = None
    if mag is not None and mag.ndim == 1:
        mag = _expand_scalar_magnetization(mag, fvec)
    nodes = surface = None
    if prop is not None and not override and index.size:
        nodes = mesh_nodes(prisms, prop, fvec=fvec)
        surface = relief_surface(prisms, prop, bounds=bounds if index.size == total else None)
    elif index.size == total:
        nodes = mesh_geometry(prisms)          # Jacobian / override models: geometry only
    flat = FlatModel(bounds, dens, mag, index, total, nodes, surface)
    if _is_prism_mesh(prisms) and index.size:
        # a regular mesh: the range guard's inputs follow from its bounds and cell size
        flat.extent_hint = (max(abs(float(v)) for v in prisms.bounds),
                            min(abs(float(d)) for d in prisms.dims))
    return flat


def _expand_scalar_magnetization(values, fvec):
    if fvec is None:
        raise TypeError("the 'magnetization' property must be a 3-component vector")
    fx, fy, fz = fvec
    values = np.asarray(values, dtype=np.float64)
    return np.stack([values * fx, values * fy, values * fz], axis=1)


def _flatten_generic(prisms, prop, override, fvec):
    rows, dens, mag, index = [], [], [], []
    total = 0
    for pos, p in enumerate(prisms):
        total = pos + 1
        if p is None:
            continue
        if prop is not None and not override and prop not in p.props:
            continue
        rows.append((float(p.x1), float(p.x2), float(p.y1), float(p.y2),
                     float(p.z1), float(p.z2)))
        index.append(pos)
        if override or prop is None:
            continue
        value = p.props[prop]
        if prop == 'density':
            dens.append(float(value))
        elif _is_scalar_property(value):
            if fvec is None:
                raise TypeError("the 'magnetization' property must be a 3-component vector")
            mag.append((value * fvec[0], value * fvec[1], value * fvec[2]))
        else:
            mx, my, mz = value
            mag.append((float(mx), float(my), float(mz)))
    arr = np.array(rows, dtype=np.float64).reshape(-1, 6)
    bounds = [np.ascontiguousarray(arr[:, c]) for c in range(6)]
    return FlatModel(bounds,
                     np.array(dens, dtype=np.float64) if (prop == 'density' and not override) else None,
                     (np.array(mag, dtype=np.float64).reshape(-1, 3)
                      if (prop == 'magnetization' and not override) else None),
                     np.array(index, dtype=np.int64), total)
