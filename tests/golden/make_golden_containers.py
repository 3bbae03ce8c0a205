"""
Golden fixture for the containers either side of the path that are not exercised by the
field fixtures: PointGrid (the equivalent layer's sources, mesh.py:186-388) and
PrismMesh.dump (UBC-GIF text, mesh.py:831-891), from the REAL reference.

Run in the build container only:  python tests/golden/make_golden_containers.py
(imports the reference like make_golden.py does; nothing of this repo's product package).
"""
import io
import os

import numpy as np

from make_golden import import_reference

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    prism, geometry, mesh, utils, gridder = import_reference()
    out = {}
    # ---- PointGrid: arrays, item access, split --------------------------------------
    area, shape = (-300.0, 900.0, 50.0, 1250.0), (6, 4)
    z = np.linspace(100, 330, 24)
    g = mesh.PointGrid(area, z, shape)
    g.addprop('density', np.arange(24) * 1.5 - 7)
    out.update(pg_area=np.array(area), pg_shape=np.array(shape), pg_zin=z, pg_x=g.x, pg_y=g.y, pg_z=g.z,
               pg_radius=np.array(g.radius), pg_spacing=np.array([g.dx, g.dy]),
               pg_density=np.asarray(g.props['density'], dtype=float))
    items = [g[i] for i in (0, 5, -1)]
    out['pg_items'] = np.array([[s.x, s.y, s.z, s.radius, s.props['density']] for s in items])
    parts = g.split((3, 2))
    out['pg_split_n'] = np.array(len(parts))
    for k, s in enumerate(parts):
        out['pg_split_%d' % k] = np.array([s.x, s.y, s.z, np.asarray(s.props['density'], dtype=float)])
        out['pg_split_area_%d' % k] = np.array(s.area, dtype=float)
    flat = mesh.PointGrid((0, 10, 2, 6), 200, (2, 3))
    out['pg_flat'] = np.array([flat.x, flat.y, flat.z])
    # ---- PrismMesh.dump: with and without masked cells ------------------------------
    m = mesh.PrismMesh((-50.0, 150.0, 10.0, 310.0, -20.0, 100.0), (3, 4, 5))
    rs = np.random.RandomState(3)
    m.addprop('density', rs.uniform(-500, 500, m.size))
    for tag, mask in (('plain', []), ('masked', [0, 7, 33, 59])):
        m.mask = list(mask)
        mf, pf = io.StringIO(), io.StringIO()
        m.dump(mf, pf, 'density')
        out['dump_%s_mesh' % tag] = np.frombuffer(mf.getvalue().encode(), dtype=np.uint8)
        out['dump_%s_prop' % tag] = np.frombuffer(pf.getvalue().encode(), dtype=np.uint8)
    out.update(dump_bounds=np.array(m.bounds), dump_shape=np.array(m.shape),
               dump_density=np.asarray(m.props['density']), dump_mask=np.array([0, 7, 33, 59]))
    np.savez_compressed(os.path.join(HERE, 'containers.npz'), **out)
    print('wrote containers.npz with', len(out), 'arrays')


if __name__ == '__main__':
    main()
