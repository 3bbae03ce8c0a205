"""
Forward modelling of the gravity fields of right rectangular prisms -- the reference's
cookbook/gravmag_grav_prism.py (BASELINE config 1) on the B200, without the plotting.

    python examples/grav_prism.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fatiando_b200 import gridder, mesher, prism   # noqa: E402

model = [mesher.Prism(-4000, -3000, -4000, -3000, 0, 2000, {'density': 1000}),
         mesher.Prism(-1000, 1000, -1000, 1000, 0, 2000, {'density': -900}),
         mesher.Prism(2000, 4000, 3000, 4000, 0, 2000, {'density': 1300})]
shape = (100, 100)
xp, yp, zp = gridder.regular((-5000, 5000, -5000, 5000), shape, z=-150)
# one call per field, exactly like the reference ...
fields = [prism.potential(xp, yp, zp, model), prism.gx(xp, yp, zp, model), prism.gy(xp, yp, zp, model),
          prism.gz(xp, yp, zp, model), prism.gxx(xp, yp, zp, model), prism.gxy(xp, yp, zp, model),
          prism.gxz(xp, yp, zp, model), prism.gyy(xp, yp, zp, model), prism.gyz(xp, yp, zp, model),
          prism.gzz(xp, yp, zp, model)]
# ... or all of them in one fused pass
fused = prism.fields(xp, yp, zp, model, prism.GRAVITY_FIELDS)
for name, one in zip(prism.GRAVITY_FIELDS, fields):
    print("%-10s min %12.5g  max %12.5g  (fused pass differs by %.1e)" % (
        name, one.min(), one.max(), abs(one - fused[name]).max()))
