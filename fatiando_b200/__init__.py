"""
fatiando_b200 -- the right-rectangular-prism potential-field forward model of
fatiando/fatiando, rebuilt for NVIDIA B200 (sm_100a).

Scope (one hot path, nothing else): ``fatiando.gravmag.prism`` and its native
kernel ``_prism``, the prism meshes that feed it, and the dense sensitivity
(Jacobian) build on top.  See DESIGN.md.

* :mod:`fatiando_b200.prism`   -- drop-in for ``fatiando.gravmag.prism``
* :mod:`fatiando_b200._prism`  -- drop-in for the Cython module ``_prism``
* :mod:`fatiando_b200.mesher`  -- ``Prism``, ``PrismMesh``, ``PrismRelief``
* :mod:`fatiando_b200.flatten` -- mesh -> SoA arrays
* :mod:`fatiando_b200.partition` -- observation / prism sharding over GPUs
* :mod:`fatiando_b200.normal`  -- Gauss-Newton products (J^T W J, J^T W r, J p) on a device-resident J
* :mod:`fatiando_b200.harvester` -- drop-in for ``fatiando.gravmag.harvester`` (planting
  inversion) on a device-resident effect engine
* :mod:`fatiando_b200.sphere`  -- drop-in for ``fatiando.gravmag.sphere`` (sibling kernel)
* :mod:`fatiando_b200.gridder`, :mod:`~fatiando_b200.utils`,
  :mod:`~fatiando_b200.constants` -- the support functions the path uses

The compute library (``libfatiando_b200.so``, C ABI in
``include/fatiando_b200.h``) is loaded lazily on first use; there is no CPU
implementation behind it.
"""
from . import constants, gridder, mesher, utils          # noqa: F401
from . import flatten, prism, _prism, partition          # noqa: F401
from . import harvester, sphere, polyprism, imaging, normal   # noqa: F401
from .mesher import Prism, PrismMesh, PrismRelief, Sphere, PolygonalPrism    # noqa: F401
from ._lib import pinned_empty                            # noqa: F401

__version__ = "0.1.0"
