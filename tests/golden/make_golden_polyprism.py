"""
Golden vectors of the polygonal-prism path from the REAL reference
(fatiando/gravmag/polyprism.py).

    python tests/golden/make_golden_polyprism.py

Case 'reg' repeats the recipe of the reference's own regression file
(fatiando/gravmag/tests/test_polyprism.py:39-68; a 30-vertex ellipse) and is checked here
against that file (tests/data/polyprism.npz); case 'many' is a seeded set of polygons with
None entries and missing properties; case 'box' is the 4-vertex prism of
tests/test_polyprism_vs_prism.py with the rectangular-prism reference's results next to it.
Nothing of this repo's product package is imported.
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
    sys.modules['fatiando'].utils = utils
    importlib.import_module('fatiando._our_duecredit')
    poly = importlib.import_module('fatiando.gravmag.polyprism')
    PolygonalPrism, Prism = geometry.PolygonalPrism, geometry.Prism
    out = {}

    def run(model, x, y, z, inc, dec, tag, pmag=None):
        for f in FIELDS:
            if f == 'tf':
                out[tag + f] = poly.tf(x, y, z, model, inc, dec, pmag=pmag)
            elif pmag is None:
                out[tag + f] = getattr(poly, f)(x, y, z, model)

    # ---- the reference's regression recipe (30 vertices on an ellipse) ----
    t = np.linspace(0, 2 * np.pi, 30, endpoint=False)          # gridder.circular_scatter without random
    verts = np.transpose([-0 + 1000 * np.cos(t), 0 + 1000 * np.sin(t)])
    try:
        pg = importlib.import_module('fatiando.gridder.point_generation')
        verts = np.transpose(pg.circular_scatter([-1000, 1000] * 2, 30))
    except Exception as exc:                                    # pragma: no cover
        print('circular_scatter unavailable:', exc)
    verts[:, 1] *= 2.5
    props = {'density': 1000, 'magnetization': utils.ang2vec(2, 25, -10)}
    model = [PolygonalPrism(verts, 100, 500, props)]
    inc, dec = -30, 20
    x, y, z = grd.regular([-4500, 4500, -5000, 5000], (100, 91), z=-10)
    theirs = np.load(os.path.join(mg.REF, 'fatiando', 'gravmag', 'tests', 'data', 'polyprism.npz'))
    run(model, x, y, z, inc, dec, 'full_')
    worst = max(float(np.max(np.abs(out['full_' + f] - theirs[f]))) for f in FIELDS)
    print('against the reference regression file: max |diff| = %.3g' % worst)
    assert worst < 1e-6
    sel = slice(None, None, 7)
    for f in FIELDS:
        out['reg_' + f] = out.pop('full_' + f)[sel]
    for k in 'xx xy xz yy yz zz'.split():
        out['reg_kernel' + k] = getattr(poly, 'kernel' + k)(x[sel], y[sel], z[sel], model[0])
    out.update(reg_x=x[sel], reg_y=y[sel], reg_z=z[sel], reg_inc=inc, reg_dec=dec, reg_verts=verts,
               reg_z1z2=np.array([100., 500.]), reg_mag=np.asarray(props['magnetization'], float))
    # ---- several polygons, None entries, missing properties ----
    rs = np.random.RandomState(17)
    polys, model = [], []
    nv = [3, 4, 5, 7, 9, 12, 6, 5]
    kind = [3, 1, 0, 2, 3, 3, 1, 2]          # 0: None, 1: density only, 2: magnetization only, 3: both
    dens = rs.uniform(-900, 900, len(nv))
    mag = rs.uniform(-3, 3, (len(nv), 3))
    vx, vy, ptr, zz = [], [], [0], []
    for q, n in enumerate(nv):
        ang = np.sort(rs.uniform(0, 2 * np.pi, n))[::-1]        # clockwise
        rad = rs.uniform(300, 900, n)
        cx, cy = rs.uniform(-2500, 2500, 2)
        px, py = cx + rad * np.cos(ang), cy + rad * np.sin(ang)
        z1 = rs.uniform(50, 500)
        z2 = z1 + rs.uniform(100, 800)
        vx += list(px); vy += list(py); ptr.append(ptr[-1] + n); zz.append((z1, z2))
        if kind[q] == 0:
            model.append(None)
            continue
        p = {}
        if kind[q] in (1, 3):
            p['density'] = float(dens[q])
        if kind[q] in (2, 3):
            p['magnetization'] = mag[q].copy()
        model.append(PolygonalPrism(np.transpose([px, py]), z1, z2, p))
    x, y, z = grd.scatter([-4000, 4000, -4000, 4000], 250, z=-30, seed=5)
    inc, dec = 55, -12
    run(model, x, y, z, inc, dec, 'many_')
    run(model, x, y, z, inc, dec, 'many_pmag_', pmag=[0.5, -1.5, 2.5])
    out.update(many_x=x, many_y=y, many_z=z, many_inc=inc, many_dec=dec, many_vx=np.array(vx),
               many_vy=np.array(vy), many_ptr=np.array(ptr), many_zz=np.array(zz), many_dens=dens,
               many_mag=mag, many_kind=np.array(kind))
    # ---- the box of test_polyprism_vs_prism.py: both references side by side ----
    props = {'density': 2670., 'magnetization': utils.ang2vec(3, 40, -20)}
    box_xy = [[-500, -700], [-500, 900], [600, 900], [600, -700]]      # clockwise, x north / y east
    pp = [PolygonalPrism(box_xy, 120, 800, props)]
    rp = [Prism(-500, 600, -700, 900, 120, 800, props)]
    x, y, z = grd.scatter([-3000, 3000, -3000, 3000], 200, z=-50, seed=9)
    inc, dec = -15, 60
    for f in FIELDS:
        if f == 'tf':
            out['box_poly_' + f] = poly.tf(x, y, z, pp, inc, dec)
            out['box_prism_' + f] = prism.tf(x, y, z, rp, inc, dec)
        else:
            out['box_poly_' + f] = getattr(poly, f)(x, y, z, pp)
            out['box_prism_' + f] = getattr(prism, f)(x, y, z, rp)
    out.update(box_x=x, box_y=y, box_z=z, box_inc=inc, box_dec=dec, box_verts=np.array(box_xy, float),
               box_z1z2=np.array([120., 800.]), box_mag=np.asarray(props['magnetization'], float))
    np.savez_compressed(os.path.join(HERE, 'polyprism.npz'), **out)
    print('polyprism.npz written:', len(out), 'arrays')


if __name__ == "__main__":
    main()
