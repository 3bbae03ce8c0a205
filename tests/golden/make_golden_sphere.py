"""
Golden vectors of the sphere path from the REAL reference (fatiando/gravmag/sphere.py).

    python tests/golden/make_golden_sphere.py

Case 'regression' repeats the recipe of the reference's own regression file
(fatiando/gravmag/tests/test_sphere.py:39-65) and is checked here against that file
(fatiando/gravmag/tests/data/sphere.npz); case 'many' is a seeded cloud of spheres with
None entries, missing properties and the dens= / pmag= overrides.  Inputs are stored next to
outputs.  Nothing of this repo's product package is imported.
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg   # noqa: E402

FIELDS = 'gz gxx gxy gxz gyy gyz gzz bx by bz tf'.split()


def main():
    prism, geometry, mesh, utils, grd = mg.import_reference()
    fat = sys.modules['fatiando']
    fat.utils = utils
    due = importlib.import_module('fatiando._our_duecredit')
    sphere = importlib.import_module('fatiando.gravmag.sphere')
    Sphere = geometry.Sphere
    out = {}
    # ---- the reference's regression recipe ----
    props = {'density': 1, 'magnetization': utils.ang2vec(1, 25, -10)}
    model = [Sphere(x=-100, y=200, z=500, radius=400, props=props)]
    inc, dec = -30, 50
    x, y, z = grd.regular([-2000, 2000, -1900, 1900], (101, 96), z=-1)
    theirs = np.load(os.path.join(mg.REF, 'fatiando', 'gravmag', 'tests', 'data', 'sphere.npz'))
    for f in FIELDS:
        got = sphere.tf(x, y, z, model, inc, dec) if f == 'tf' else getattr(sphere, f)(x, y, z, model)
        np.testing.assert_allclose(got, theirs[f], rtol=0, atol=1e-10)      # their own tolerance
        out['reg_' + f] = got[::4]            # every 4th point is stored (the full grid was checked above)
    for k in 'xx xy xz yy yz zz'.split():
        out['reg_kernel' + k] = getattr(sphere, 'kernel' + k)(x, y, z, model[0])[::4]
    out.update(reg_x=x[::4], reg_y=y[::4], reg_z=z[::4], reg_inc=inc, reg_dec=dec,
               reg_sphere=np.array([-100., 200., 500., 400.]), reg_mag=np.asarray(props['magnetization'], float))
    # ---- a cloud of spheres ----
    rs = np.random.RandomState(11)
    m = 40
    geo = np.column_stack([rs.uniform(-3000, 3000, m), rs.uniform(-3000, 3000, m), rs.uniform(200, 2500, m),
                           rs.uniform(50, 400, m)])
    dens = rs.uniform(-800, 800, m)
    mag = rs.uniform(-3, 3, (m, 3))
    kind = rs.randint(0, 4, m)       # 0: None, 1: density only, 2: magnetization only, 3: both
    model = []
    for q in range(m):
        if kind[q] == 0:
            model.append(None)
            continue
        p = {}
        if kind[q] in (1, 3):
            p['density'] = float(dens[q])
        if kind[q] in (2, 3):
            p['magnetization'] = mag[q].copy()
        model.append(Sphere(*geo[q], props=p))
    x, y, z = grd.scatter([-4000, 4000, -4000, 4000], 300, z=-150, seed=2)
    inc, dec = 40, -25
    for f in FIELDS:
        if f == 'tf':
            out['many_' + f] = sphere.tf(x, y, z, model, inc, dec)
            out['many_pmag_' + f] = sphere.tf(x, y, z, model, inc, dec, pmag=[1.5, -0.5, 2.0])
        elif f in ('bx', 'by', 'bz'):
            out['many_' + f] = getattr(sphere, f)(x, y, z, model)
            out['many_pmag_' + f] = getattr(sphere, f)(x, y, z, model, pmag=[1.5, -0.5, 2.0])
        else:
            out['many_' + f] = getattr(sphere, f)(x, y, z, model)
            out['many_dens_' + f] = getattr(sphere, f)(x, y, z, model, dens=-250.)
    out.update(many_x=x, many_y=y, many_z=z, many_inc=inc, many_dec=dec, many_geo=geo, many_dens=dens,
               many_mag=mag, many_kind=kind)
    np.savez_compressed(os.path.join(HERE, 'sphere.npz'), **out)
    print('sphere.npz written:', len(out), 'arrays')


if __name__ == "__main__":
    main()
