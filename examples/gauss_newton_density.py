"""
One Gauss-Newton step of a linear density inversion with the sensitivity matrix kept on the GPU
(what fatiando.inversion does with numpy on a host J: Misfit.hessian = 2 J^T J, misfit.py:250-256;
Misfit.gradient = -2 J^T r, misfit.py:280-294; optimization.linear solves H p = -g with Jacobi
preconditioning, optimization.py:51-97; damping as in regularization.Damping).

    python examples/gauss_newton_density.py

gz of a 10 x 10 x 4 mesh observed on a 40 x 40 grid; the step from p = 0 with a small damping
recovers a model that reproduces the data.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fatiando_b200 import gridder, mesher, normal, prism   # noqa: E402

mesh = mesher.PrismMesh((0, 5000, 0, 5000, 0, 2000), (4, 10, 10))
true = np.zeros(mesh.size)
true[[144, 145, 154, 155, 244, 245, 254, 255]] = 600.0
mesh.addprop('density', true)
x, y, z = gridder.regular((0, 5000, 0, 5000), (40, 40), z=-100)
data = prism.gz(x, y, z, mesh)

t0 = time.perf_counter()
with prism.Model(mesh, None, keep_all=True) as model, normal.SensitivityBlock(model, x, y, z, 'gz') as J:
    hessian = J.gram(factor=2.0)                    # 2 J^T J           (fp64 SYRK on the GPU)
    gradient = J.tdot(data, factor=-2.0)            # -2 J^T (d - J 0)
    mu = 1e-10 * np.trace(hessian) / mesh.size      # damping
    hessian[np.diag_indices_from(hessian)] += 2.0 * mu
    diag = np.abs(np.diag(hessian))                 # Jacobi preconditioning, optimization.py:90-95
    diag[diag < 1e-10] = 1e-10
    estimate = np.linalg.solve(hessian / diag[:, None], -gradient / diag)
    predicted = J.dot(estimate)                     # J p, still on the GPU matrix
print("J %d x %d on the GPU, normal equations solved in %.2f s" % (J.shape + (time.perf_counter() - t0,)))
print("data misfit: max |d - J p| = %.3g mGal of max |d| = %.3g mGal" % (np.abs(data - predicted).max(),
                                                                         np.abs(data).max()))
mesh.addprop('density', estimate)
again = prism.gz(x, y, z, mesh)                     # the forward model of the estimate = J p
assert np.allclose(again, predicted, rtol=1e-9, atol=1e-10 * np.abs(data).max())
assert np.abs(data - predicted).max() < 1e-3 * np.abs(data).max()
print("ok")
