"""
Gravity imaging by migration on the B200 -- ``fatiando.gravmag.imaging.migrate``
(reference imaging.py:64-120, Zhdanov et al. 2011).

The reference builds, layer by layer, the transposed sensitivity of every mesh cell with
one ``prism.gz(x, y, z, [cell], dens=1)`` call per cell and multiplies it with the data
(imaging.py:116-118).  Here that product is the on-the-fly adjoint ``J^T gz`` of the whole
mesh (:func:`fatiando_b200.prism.adjoint`, node-sharing kernel): nothing of size
``ncells x ndata`` is ever stored.  ``sandwich`` and ``geninv`` of the reference are
FFT-based and do not touch the prism kernels; they are not carried.
"""
import numpy

from . import prism
from .constants import G
from .mesher import PrismMesh


def _makemesh(x, y, shape, zmin, zmax, nlayers):
    """A prism mesh bounded by the data (imaging.py:260-267)."""
    ny, nx = shape
    return PrismMesh([x.min(), x.max(), y.min(), y.max(), zmin, zmax], (nlayers, ny, nx))


def migrate(x, y, z, gz, zmin, zmax, meshshape, power=0.5, scale=1, device=0):
    """
    3D potential-field migration of gravity data (gz, in mGal).

    * x, y : 1D-arrays, coordinates of the data points (need not be a regular grid)
    * z : float or 1D-array, their z coordinate
    * gz : 1D-array, the gravity anomaly
    * zmin, zmax : top and bottom of the region where the density is estimated
    * meshshape = (nz, ny, nx) : number of prisms of the output mesh
    * power : power law of the depth weights ``|depth|**power / (2 G sqrt(pi))``
    * scale : scale factor of the depth weights

    Returns the :class:`~fatiando_b200.mesher.PrismMesh` with the estimate in
    ``mesh.props['density']``.
    """
    nlayers, ny, nx = meshshape
    x = numpy.ascontiguousarray(x, dtype=numpy.float64)
    y = numpy.ascontiguousarray(y, dtype=numpy.float64)
    gz = numpy.ascontiguousarray(gz, dtype=numpy.float64)
    mesh = _makemesh(x, y, (ny, nx), zmin, zmax, nlayers)
    z = numpy.ascontiguousarray(z * numpy.ones_like(x))
    dx, dy, dz = mesh.dims
    depths = mesh.get_zs()[:-1] + 0.5 * dz
    weights = numpy.abs(depths) ** power / (2 * G * numpy.sqrt(numpy.pi))
    adj = prism.adjoint(x, y, z, mesh, gz, 'gz', device=device)      # J^T gz, all layers at once
    density = (scale * weights)[:, None] * adj.reshape(nlayers, ny * nx)
    mesh.addprop('density', density.ravel())
    return mesh
