"""
Observation-point generators used by the prism configs
(reference: fatiando/gridder/point_generation.py:20-96, gridder/utils.py:6-38).
"""
import numpy


def regular(area, shape, z=None):
    """
    Regular grid of points; x varies slowest.

    * area = (x1, x2, y1, y2), shape = (nx, ny); returns ``[x, y]`` or
      ``[x, y, z]`` as raveled 1-D arrays, in the reference's node order
      (point_generation.py:89-96).
    """
    nx, ny = shape
    x1, x2, y1, y2 = area
    if x1 > x2 or y1 > y2:
        raise ValueError("Invalid area dimensions {}".format(area))
    xs = numpy.linspace(x1, x2, nx)
    ys = numpy.linspace(y1, y2, ny)
    # the same construction as the reference (not an equivalent one): the grids
    # must be the very same doubles
    arrays = list(numpy.meshgrid(ys, xs))[::-1]
    if z is not None:
        arrays.append(z * numpy.ones(nx * ny, dtype=float))
    return [i.ravel() for i in arrays]


def scatter(area, n, z=None, seed=None):
    """
    n random points uniformly distributed in area (point_generation.py:99-150).
    """
    x1, x2, y1, y2 = area
    if x1 > x2 or y1 > y2:
        raise ValueError("Invalid area dimensions {}".format(area))
    numpy.random.seed(seed)
    arrays = [numpy.random.uniform(x1, x2, n), numpy.random.uniform(y1, y2, n)]
    if z is not None:
        arrays.append(z * numpy.ones(n))
    return arrays


def spacing(area, shape):
    """Spacing ``[dx, dy]`` between the nodes of a regular grid."""
    x1, x2, y1, y2 = area
    nx, ny = shape
    return [(x2 - x1) / (nx - 1), (y2 - y1) / (ny - 1)]
