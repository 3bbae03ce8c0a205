"""
Golden fixtures of the harvester path from the REAL reference.

    python tests/golden/make_golden_harvester.py

Imports fatiando/gravmag/harvester.py from /root/reference with the shims of
make_golden.py and runs harvester.harvest / iharvest on small synthetic
problems; stores inputs, the accretion sequence, goal/misfit per accretion,
the estimate and the (float32) predicted data.  Nothing of this repo's product
package is imported.
"""
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg   # noqa: E402


def reference_modules():
    prism, geometry, mesh, utils, grd = mg.import_reference()
    m = sys.modules['fatiando.mesher']
    m.Prism, m.Tesseroid, m.PrismMesh = geometry.Prism, geometry.Tesseroid, mesh.PrismMesh
    sys.modules['fatiando'].utils = utils
    hv = importlib.import_module('fatiando.gravmag.harvester')
    return hv, prism, geometry, mesh, utils, grd


def run_case(hv, mesh, data, locations, compactness, threshold, restrict=None):
    seeds = hv.sow(locations, mesh)
    restrict_l = hv._test_restriction(restrict)
    order, goals, misfits, regs = [], [], [], []
    for update in hv.iharvest(data, seeds, mesh, compactness, threshold, restrict_l):
        estimate, predicted, new = update[:3]
        order.append(-1 if new is None else int(new.i))
        goals.append(float(update[4]))
        misfits.append(float(update[5]))
        regs.append(float(update[6]))
    est = hv.fmt_estimate(estimate, mesh.size)
    out = {'seed_index': np.array([s.i for s in seeds]), 'order': np.array(order),
           'goal': np.array(goals), 'misfit': np.array(misfits), 'regularizer': np.array(regs)}
    for k in est:
        out['estimate_' + k] = np.array([np.asarray(est[k][i], dtype=float) for i in range(mesh.size)])
    for i, p in enumerate(predicted):
        out['predicted_%d' % i] = np.asarray(p)
    return out


def main():
    hv, prism, geometry, meshmod, utils, grd = reference_modules()
    Prism, PrismMesh = geometry.Prism, meshmod.PrismMesh
    rs = np.random.RandomState(5)
    # ---- case A: gz + gzz, two bodies of different density, two seeds each -----
    bounds = (0, 3000, 0, 3000, 0, 1500)
    model = [Prism(600, 1400, 700, 1500, 200, 800, {'density': 800}),
             Prism(1900, 2500, 1600, 2400, 300, 900, {'density': -500})]
    shape = (16, 16)
    x, y, z = grd.regular(bounds[:4], shape, z=-100)
    gz = prism.gz(x, y, z, model) + rs.normal(0, 0.02, x.size)
    gzz = prism.gzz(x, y, z, model) + rs.normal(0, 0.5, x.size)
    mesh = PrismMesh(bounds, (6, 12, 12))
    locs = [[1000, 1100, 500, {'density': 800}], [2200, 2000, 600, {'density': -500}]]
    w = hv.weights(x, y, hv.sow(locs, mesh), [1500, 1500])
    data = [hv.Gz(x, y, z, gz), hv.Gzz(x, y, z, gzz, weights=w)]
    out = run_case(hv, mesh, data, locs, compactness=0.5, threshold=0.0005)
    out.update(x=x, y=y, z=z, gz=gz, gzz=gzz, weights=w, mesh_bounds=np.array(bounds, float),
               mesh_shape=np.array([6, 12, 12]), locations=json.dumps(locs),
               compactness=0.5, threshold=0.0005)
    np.savez_compressed(os.path.join(HERE, 'harvester_grav.npz'), **out)
    print('grav: %d accretions' % (len(out['order']) - 1))
    # ---- case B: total-field anomaly, induced magnetization, carved mesh ----------
    inc, dec = -40.0, 12.0
    model = [Prism(800, 1800, 900, 2100, 150, 650, {'magnetization': 5.0})]
    x, y, z = grd.regular(bounds[:4], (14, 15), z=-80)
    tf = prism.tf(x, y, z, model, inc, dec) + rs.normal(0, 1.0, x.size)
    mesh = PrismMesh(bounds, (5, 10, 10))
    mesh.mask = [0, 1, 2, 10, 11, 55]
    locs = [[1300, 1500, 400, {'magnetization': 5.0}]]
    data = [hv.TotalField(x, y, z, tf, inc, dec)]
    out = run_case(hv, mesh, data, locs, compactness=0.1, threshold=0.001, restrict=['above'])
    out.update(x=x, y=y, z=z, tf=tf, inc=inc, dec=dec, mesh_bounds=np.array(bounds, float),
               mesh_shape=np.array([5, 10, 10]), mask=np.array(mesh.mask), locations=json.dumps(locs),
               compactness=0.1, threshold=0.001, restrict=json.dumps(['above']))
    np.savez_compressed(os.path.join(HERE, 'harvester_mag.npz'), **out)
    print('mag: %d accretions' % (len(out['order']) - 1))
    # ---- case C: joint gz + total field, one seed carrying both properties, one only density ----
    inc, dec = 60.0, -8.0
    model = [Prism(900, 1700, 800, 1800, 200, 700, {'density': 600, 'magnetization': utils.ang2vec(4, inc, dec)}),
             Prism(2000, 2600, 1900, 2500, 250, 650, {'density': -400})]
    x, y, z = grd.regular(bounds[:4], (15, 14), z=-60)
    gz = prism.gz(x, y, z, model) + rs.normal(0, 0.01, x.size)
    tf = prism.tf(x, y, z, model, inc, dec) + rs.normal(0, 0.5, x.size)
    mesh = PrismMesh(bounds, (5, 10, 10))
    locs = [[1300, 1300, 450, {'density': 600, 'magnetization': 4.0}],
            [2300, 2200, 450, {'density': -400}]]
    data = [hv.Gz(x, y, z, gz), hv.TotalField(x, y, z, tf, inc, dec, weights=0.7)]
    out = run_case(hv, mesh, data, locs, compactness=0.3, threshold=0.0008)
    out.update(x=x, y=y, z=z, gz=gz, tf=tf, inc=inc, dec=dec, mesh_bounds=np.array(bounds, float),
               mesh_shape=np.array([5, 10, 10]), locations=json.dumps(locs), compactness=0.3,
               threshold=0.0008)
    np.savez_compressed(os.path.join(HERE, 'harvester_joint.npz'), **out)
    print('joint: %d accretions' % (len(out['order']) - 1))


if __name__ == "__main__" and "--imaging" not in sys.argv:
    main()


def imaging_case():
    """fatiando.gravmag.imaging.migrate on a small synthetic (SURVEY 8(f) row 2's caller)."""
    hv, prism, geometry, meshmod, utils, grd = reference_modules()
    sys.modules['fatiando.gravmag'].transform = type(sys)('transform_stub')
    sys.modules['fatiando.gravmag.transform'] = sys.modules['fatiando.gravmag'].transform
    sys.modules['fatiando.gravmag'].prism = prism
    imaging = importlib.import_module('fatiando.gravmag.imaging')
    model = [geometry.Prism(600, 1400, 700, 1500, 200, 800, {'density': 800})]
    x, y, z = grd.regular((0, 3000, 0, 3000), (21, 19), z=-20)
    gz = prism.gz(x, y, z, model)
    mesh = imaging.migrate(x, y, z, gz, 0, 1200, (6, 9, 10), power=0.7, scale=2.0)
    np.savez_compressed(os.path.join(HERE, 'imaging_migrate.npz'), x=x, y=y, z=z, gz=gz,
                        density=np.asarray(mesh.props['density'], dtype=float),
                        bounds=np.asarray(mesh.bounds, dtype=float), shape=np.array(mesh.shape))
    print('imaging: density range', float(np.min(mesh.props['density'])), float(np.max(mesh.props['density'])))


if __name__ == "__main__" and "--imaging" in sys.argv:
    imaging_case()
