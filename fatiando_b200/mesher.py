"""
Containers of the forward-model path: ``Prism``, ``PrismMesh``, ``PrismRelief`` and, for the
sibling kernels, ``Sphere``, ``PolygonalPrism`` and ``PointGrid`` (the equivalent layer's
grid of point sources, mesh.py:186-388).

They behave like the reference's classes (fatiando/mesher/geometry.py:170-260,
fatiando/mesher/mesh.py:391-492 and :495-665): iterables / sequences of
``Prism``-or-``None`` with a ``props`` dictionary, same cell ordering, same
coordinate expressions (so the bounds are the very same doubles).  The forward
model never iterates them cell by cell, though: :mod:`fatiando_b200.flatten`
turns a whole mesh into SoA arrays with vectorised, bit-identical expressions.
Objects of the reference's own classes are accepted everywhere these are.
"""
import copy as _copy

import numpy as np


class Prism(object):
    """
    A right rectangular prism (x -> North, y -> East, z -> Down).

    * x1, x2 : south and north borders
    * y1, y2 : west and east borders
    * z1, z2 : top and bottom
    * props : dict of physical properties, e.g. ``{'density': 200}`` or
      ``{'magnetization': [mx, my, mz]}``

    >>> p = Prism(1, 2, 3, 4, 5, 6, {'density': 200})
    >>> p.get_bounds()
    [1.0, 2.0, 3.0, 4.0, 5.0, 6.0]
    >>> print(p)
    x1:1 | x2:2 | y1:3 | y2:4 | z1:5 | z2:6 | density:200
    """

    def __init__(self, x1, x2, y1, y2, z1, z2, props=None):
        self.props = dict(props) if props is not None else {}
        self.x1, self.x2 = float(x1), float(x2)
        self.y1, self.y2 = float(y1), float(y2)
        self.z1, self.z2 = float(z1), float(z2)

    def addprop(self, prop, value):
        """Set (or overwrite) a physical property."""
        self.props[prop] = value

    def get_bounds(self):
        """``[x1, x2, y1, y2, z1, z2]``"""
        return [self.x1, self.x2, self.y1, self.y2, self.z1, self.z2]

    def center(self):
        """Coordinates of the centre of the prism."""
        return np.array([0.5 * (self.x1 + self.x2), 0.5 * (self.y1 + self.y2),
                         0.5 * (self.z1 + self.z2)])

    def copy(self):
        return _copy.deepcopy(self)

    def __str__(self):
        items = list(zip(('x1', 'x2', 'y1', 'y2', 'z1', 'z2'), self.get_bounds()))
        items += [(k, self.props[k]) for k in sorted(self.props)]
        return ' | '.join('%s:%g' % (k, v) for k, v in items)


class Sphere(object):
    """
    A sphere.  x -> North, y -> East, z -> Down.

    * x, y, z : coordinates of the centre
    * radius : radius of the sphere
    * props : dict of physical properties, e.g.
      ``{'density': 2670, 'magnetization': [mx, my, mz]}``

    >>> s = Sphere(1, 2, 3, 10, {'magnetization': 200})
    >>> s.props['magnetization']
    200
    >>> s.addprop('density', 20)
    >>> print(s)
    x:1 | y:2 | z:3 | radius:10 | density:20 | magnetization:200
    """

    def __init__(self, x, y, z, radius, props=None):
        self.props = props if props is not None else {}
        self.x, self.y, self.z = float(x), float(y), float(z)
        self.radius = float(radius)
        self.center = np.array([x, y, z])

    def addprop(self, prop, value):
        self.props[prop] = value

    def copy(self):
        return Sphere(self.x, self.y, self.z, self.radius, props=dict(self.props))

    def __str__(self):
        names = [('x', self.x), ('y', self.y), ('z', self.z), ('radius', self.radius)]
        names.extend((p, self.props[p]) for p in sorted(self.props))
        return ' | '.join('%s:%g' % (n, v) for n, v in names)


class PolygonalPrism(object):
    """
    A prism with a polygonal horizontal cross-section.  x -> North, y -> East, z -> Down;
    ``vertices`` = list of ``[x, y]`` pairs, **clockwise** (the other way round flips the
    sign, like the reference, geometry.py:502-546); ``z1, z2`` = top and bottom.

    >>> p = PolygonalPrism([[0, 0], [0, 2], [1, 2], [1, 0]], 0, 3, {'density': 50})
    >>> p.nverts, p.z2
    (4, 3.0)
    """

    def __init__(self, vertices, z1, z2, props=None):
        self.props = props if props is not None else {}
        self.x = np.array([v[0] for v in vertices], dtype=np.float64)
        self.y = np.array([v[1] for v in vertices], dtype=np.float64)
        self.z1, self.z2 = float(z1), float(z2)
        self.nverts = len(vertices)

    def addprop(self, prop, value):
        self.props[prop] = value

    def copy(self):
        return PolygonalPrism(np.transpose([self.x, self.y]), self.z1, self.z2, props=dict(self.props))


class PointGrid(object):
    """
    A regular grid of point sources: spheres of unit volume (mesh.py:186-388), what the
    equivalent layer is made of (eqlayer.py:100-111).  A sequence of :class:`Sphere`, rows
    along x (north), columns along y; ``z`` a number or one value per point.

    >>> g = PointGrid([0, 10, 2, 6], 200, (2, 3))
    >>> g.shape, g.size, (g.dx, g.dy)
    ((2, 3), 6, (10.0, 2.0))
    >>> [list(map(float, p.center)) for p in g][1::4]
    [[0.0, 4.0, 200.0], [10.0, 6.0, 200.0]]
    >>> z = np.linspace(0, 1100, 12)
    >>> g = PointGrid((0, 3, 0, 2), z, (4, 3))
    >>> g.addprop('bla', np.arange(1, 13))
    >>> [(s.props['bla'].tolist(), s.x.tolist(), s.y.tolist(), s.z.tolist()) for s in g.split((2, 3))][4]
    ([8, 11], [2.0, 3.0], [1.0, 1.0], [700.0, 1000.0])
    """

    def __init__(self, area, z, shape, props=None):
        from . import gridder
        self.area = area
        self.shape = shape
        self.props = props if props is not None else {}
        nx, ny = shape
        self.size = nx * ny
        self.z = np.zeros(self.size) + z
        self.radius = float(np.cbrt(3. / (4. * np.pi)))          # unit volume
        self.x, self.y = gridder.regular(area, shape)
        with np.errstate(all='ignore'):        # a one-point row or column has no spacing (nan), like the reference
            self.dx, self.dy = gridder.spacing(area, shape)
        self.i = 0

    def __len__(self):
        return self.size

    def __getitem__(self, index):
        if not isinstance(index, (int, np.integer)):
            raise IndexError('Invalid index type. Should be int.')
        if index >= self.size or index < -self.size:
            raise IndexError('Grid index out of range.')
        if index < 0:
            index += self.size
        return Sphere(self.x[index], self.y[index], self.z[index], self.radius,
                      props={p: self.props[p][index] for p in self.props})

    def __iter__(self):
        self.i = 0
        return self

    def __next__(self):
        if self.i >= self.size:
            raise StopIteration
        point = self[self.i]
        self.i += 1
        return point

    next = __next__

    def addprop(self, prop, values):
        """One value of ``prop`` per point of the grid."""
        self.props[prop] = values

    def split(self, shape):
        """
        The grid cut into ``shape = (nx, ny)`` subgrids (a list of :class:`PointGrid`, x blocks
        outermost); both counts must divide the grid's.
        """
        nx, ny = shape
        totalx, totaly = self.shape
        if totalx % nx != 0 or totaly % ny != 0:
            raise ValueError('Cannot split! nx and ny must be divisible by grid shape')
        x1, x2, y1, y2 = self.area
        mx, my = totalx // nx, totaly // ny
        xstarts = np.linspace(x1, x2, totalx)[::mx]
        ystarts = np.linspace(y1, y2, totaly)[::my]
        lx, ly = self.dx * (mx - 1), self.dy * (my - 1)
        zgrid = np.reshape(self.z, self.shape)
        pgrids = {p: np.reshape(self.props[p], self.shape) for p in self.props}
        parts = []
        for i, xs in enumerate(xstarts):
            for j, ys in enumerate(ystarts):
                block = (slice(i * mx, (i + 1) * mx), slice(j * my, (j + 1) * my))
                parts.append(PointGrid([xs, xs + lx, ys, ys + ly], zgrid[block].ravel(), (mx, my),
                                       {p: pgrids[p][block].ravel() for p in pgrids}))
        return parts

    def copy(self):
        return _copy.deepcopy(self)


class _PrismSequence(object):
    """Iteration protocol shared by the two meshes (py2 ``next`` kept too)."""

    def __len__(self):
        return self.size

    def __iter__(self):
        self.i = 0
        return self

    def __next__(self):
        if self.i >= self.size:
            raise StopIteration
        cell = self[self.i]
        self.i += 1
        return cell

    next = __next__

    def copy(self):
        return _copy.deepcopy(self)


class PrismMesh(_PrismSequence):
    """
    A regular 3-D mesh of prisms.

    Cells are ordered x fastest, then y, then z (layers), like the reference
    (mesh.py:499-507): in a ``(3, 3, 3)`` mesh, index 14 is layer 1, row 1,
    column 2.

    * bounds = [xmin, xmax, ymin, ymax, zmin, zmax]
    * shape = (nz, ny, nx), integers
    * props : dict of per-cell sequences

    ``mesh[i]`` is a :class:`Prism`, or ``None`` for a masked cell
    (see :meth:`carvetopo`).

    >>> mesh = PrismMesh((0, 1, 0, 2, 0, 3), (1, 2, 2))
    >>> for p in mesh:
    ...     print(p)
    x1:0 | x2:0.5 | y1:0 | y2:1 | z1:0 | z2:3
    x1:0.5 | x2:1 | y1:0 | y2:1 | z1:0 | z2:3
    x1:0 | x2:0.5 | y1:1 | y2:2 | z1:0 | z2:3
    x1:0.5 | x2:1 | y1:1 | y2:2 | z1:0 | z2:3
    """

    celltype = Prism

    def __init__(self, bounds, shape, props=None):
        nz, ny, nx = shape
        for v in (nx, ny, nz):
            if not isinstance(v, (int, np.integer)):
                raise AttributeError(
                    'Invalid mesh shape {}. shape must be integers'.format(str(shape)))
        x1, x2, y1, y2, z1, z2 = bounds
        self.shape = (int(nz), int(ny), int(nx))
        self.size = int(nx * ny * nz)
        # true division, like the reference (mesh.py:595-597)
        self.dims = ((x2 - x1) / nx, (y2 - y1) / ny, (z2 - z1) / nz)
        self.bounds = bounds
        self.props = props if props is not None else {}
        self.mask = []
        self.zdown = True
        self.i = 0

    def __getitem__(self, index):
        if index >= self.size or index < -self.size:
            raise IndexError('mesh index out of range')
        if index < 0:
            index += self.size
        if index in self.mask:
            return None
        nz, ny, nx = self.shape
        k, rest = divmod(index, nx * ny)
        j, i = divmod(rest, nx)
        dx, dy, dz = self.dims
        x1 = self.bounds[0] + dx * i
        y1 = self.bounds[2] + dy * j
        z1 = self.bounds[4] + dz * k
        props = {p: self.props[p][index] for p in self.props}
        return self.celltype(x1, x1 + dx, y1, y1 + dy, z1, z1 + dz, props=props)

    def addprop(self, prop, values):
        """Attach one value of ``prop`` per cell (mesh order)."""
        self.props[prop] = values

    def carvetopo(self, x, y, height, below=False):
        """
        Mask the cells above (or, with ``below=True``, below) a topography.

        ``x, y, height`` are scattered samples of the surface; it is
        interpolated (cubic) onto the cell centres.  Same decision rule as
        mesh.py:667-727.
        """
        import scipy.interpolate
        nz, ny, nx = self.shape
        x1, x2, y1, y2, z1, z2 = self.bounds
        dx, dy, dz = self.dims
        xc = (np.arange(x1, x2, dx) + 0.5 * dx)[:nx]
        yc = (np.arange(y1, y2, dy) + 0.5 * dy)[:ny]
        zc = (np.arange(z1, z2, dz) + 0.5 * dz)[:nz]
        XC, YC = np.meshgrid(xc, yc)
        topo = scipy.interpolate.griddata((x, y), height, (XC, YC), method='cubic').ravel()
        if self.zdown:
            topo = -1 * topo
        # like the reference, only a *masked* interpolation result drops a column;
        # NaN (outside the data hull) compares false and keeps the cells
        nodata = np.ma.getmaskarray(topo)
        surface = np.ma.getdata(topo)
        sign = 1 if self.zdown else -1
        for k, cellz in enumerate(zc):
            if below:
                drop = nodata | (sign * cellz > sign * surface)
            else:
                drop = nodata | (sign * cellz < sign * surface)
            self.mask.extend((k * nx * ny + np.nonzero(drop)[0]).tolist())

    def _nodes(self, axis):
        lo, hi = self.bounds[2 * axis], self.bounds[2 * axis + 1]
        step, count = self.dims[axis], self.shape[2 - axis]
        return np.arange(lo, hi + step, step)[:count + 1]

    def get_xs(self):
        """x coordinates of the cell boundaries."""
        return self._nodes(0)

    def get_ys(self):
        """y coordinates of the cell boundaries."""
        return self._nodes(1)

    def get_zs(self):
        """z coordinates of the cell boundaries."""
        return self._nodes(2)

    def get_layer(self, i):
        """List of the cells of the i-th layer."""
        nz, ny, nx = self.shape
        if i >= nz or i < 0:
            raise IndexError('Layer index %d is out of range.' % (i))
        return [self[p] for p in range(i * nx * ny, (i + 1) * nx * ny)]

    def layers(self):
        """Iterate over the layers."""
        for i in range(self.shape[0]):
            yield self.get_layer(i)

    def dump(self, meshfile, propfile, prop):
        """
        Write the mesh and one property in the text format of UBC-GIF's MeshTools3D
        (mesh.py:831-891): the mesh file holds ``ny nx nz``, the origin ``y1 x1 -z1`` and the
        cell sizes as ``count*size`` per axis; the property file one value per line with z
        fastest, then x, then y (Fortran order of the ``(nz, ny, nx)`` array), ``%.4f``,
        masked cells as -10000000.  ``meshfile`` / ``propfile``: file names or open files.

        >>> from io import StringIO
        >>> meshfile, densfile = StringIO(), StringIO()
        >>> mesh = PrismMesh((0, 10, 0, 20, 0, 5), (1, 2, 2))
        >>> mesh.addprop('density', [1, 2, 3, 4])
        >>> mesh.dump(meshfile, densfile, 'density')
        >>> print(meshfile.getvalue().strip())
        2 2 1
        0 0 0
        2*10
        2*5
        1*5
        >>> print(densfile.getvalue().strip())
        1.0000
        3.0000
        2.0000
        4.0000
        """
        if prop not in self.props:
            raise ValueError("mesh doesn't have a '%s' property." % (prop))
        nz, ny, nx = self.shape
        dx, dy, dz = self.dims
        header = ["%d %d %d" % (ny, nx, nz),
                  "%g %g %g" % (self.bounds[2], self.bounds[0], -self.bounds[4]),
                  "%d*%g" % (ny, dy), "%d*%g" % (nx, dx), "%d*%g" % (nz, dz)]
        text = "\n".join(header)              # no newline after the last line, like the reference
        if isinstance(meshfile, str):
            with open(meshfile, 'w') as f:
                f.write(text)
        else:
            meshfile.write(text)
        values = np.array([float(v) for v in self.props[prop]], dtype=np.float64)
        if len(self.mask):
            values[np.asarray(self.mask, dtype=np.int64)] = -10000000
        np.savetxt(propfile, values.reshape(self.shape).ravel(order='F'), fmt='%.4f')


class PrismRelief(_PrismSequence):
    """
    A relief (topography, basin, Moho ...) made of one prism per grid node.

    * ref : reference level; a prism spans from its node's z to ``ref``
      (top on z when z <= ref, otherwise top on ref)
    * dims = (dy, dx) : horizontal prism dimensions
    * nodes = [x, y, z] : coordinates of the centre of each prism's free face

    Physical properties added with :meth:`addprop` change sign where the node
    lies below the reference (z > ref), like mesh.py:484-488.
    """

    def __init__(self, ref, dims, nodes):
        x, y, z = nodes
        if len(x) != len(y) != len(z):
            raise ValueError("nodes has x, y, z coordinate arrays of different lengths")
        self.x, self.y, self.z = x, y, z
        self.size = len(x)
        self.ref = ref
        self.dy, self.dx = dims
        self.props = {}
        self.i = 0

    def __getitem__(self, index):
        if index < 0:
            index += self.size
        xc, yc, zc = self.x[index], self.y[index], self.z[index]
        top, bottom = (zc, self.ref) if zc <= self.ref else (self.ref, zc)
        props = {p: self.props[p][index] for p in self.props}
        return Prism(xc - 0.5 * self.dx, xc + 0.5 * self.dx, yc - 0.5 * self.dy,
                     yc + 0.5 * self.dy, top, bottom, props=props)

    def addprop(self, prop, values):
        """
        Attach one value per prism.  Where the relief is below the reference
        level the stored value has the opposite sign.
        """
        self.props[prop] = [-v if self.z[i] > self.ref else v
                            for i, v in enumerate(values)]
