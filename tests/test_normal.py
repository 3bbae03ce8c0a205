"""
Gauss-Newton products on a device-resident sensitivity matrix (fatiando_b200/normal.py,
csrc/normal.cu) against numpy on the oracle's Jacobian: J^T W J (Misfit.hessian,
misfit.py:250-256), J^T W r (Misfit.gradient, misfit.py:280-294), J p.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_jacobian(oracle, flat, xp, yp, zp, field):
    return oracle.jacobian(field, xp, yp, zp, flat.bounds, [1.0], threads=4)


@pytest.mark.parametrize("shape,npts", [((5, 9, 11), 301), ((2, 3, 4), 7), ((3, 8, 16), 16), ((1, 1, 1), 40)])
def test_gram_gradient_and_predicted_data(gpu_lib, oracle, shape, npts):
    from fatiando_b200 import prism, mesher, gridder, normal
    from fatiando_b200.flatten import flatten
    rs = np.random.RandomState(17)
    mesh = mesher.PrismMesh((0, 5000, -300, 4100, 10, 2000), shape)
    flat = flatten(mesh, None, override=True)
    xp, yp, zp = gridder.scatter((-500, 5500, -800, 4600), npts, z=-120., seed=2)
    w = rs.uniform(0.2, 3.0, npts)
    r = rs.normal(size=npts)
    p = rs.uniform(-500, 500, flat.size)
    J = _oracle_jacobian(oracle, flat, xp, yp, zp, 'gz')          # in mGal, like the reference
    with prism.Model(mesh, None, keep_all=True) as model, \
            normal.SensitivityBlock(model, xp, yp, zp, 'gz') as block:
        assert block.shape == J.shape
        Jd = block.to_host()
        assert np.max(np.abs(Jd - J)) <= 1e-9 * np.abs(J).max()
        for weights, factor in ((None, 1.0), (w, 2.0 * 0.37)):
            G = block.gram(weights, factor)
            ref = factor * (Jd.T @ (Jd if weights is None else weights[:, None] * Jd))
            # |sum of n products|: rounding of either summation order stays below n eps sum |a||b|
            floor = npts * 2.0 ** -52 * abs(factor) * (np.abs(Jd).T @ (np.abs(Jd) if weights is None
                                                                        else weights[:, None] * np.abs(Jd)))
            assert G.shape == (flat.size, flat.size)
            assert np.all(np.abs(G - ref) <= floor + 1e-300)
            assert np.array_equal(G, G.T)                       # mirrored, exactly symmetric
            g = block.tdot(r, weights, -factor)
            gref = -factor * (Jd.T @ (r if weights is None else weights * r))
            gfloor = npts * 2.0 ** -52 * abs(factor) * (np.abs(Jd).T @ np.abs(r if weights is None else weights * r))
            assert np.all(np.abs(g - gref) <= gfloor + 1e-300)
        y = block.dot(p)
        yfloor = flat.size * 2.0 ** -52 * (np.abs(Jd) @ np.abs(p))
        assert np.all(np.abs(y - Jd @ p) <= yfloor + 1e-300)
        # J p is the forward field of the densities p (linearity)
        mesh.addprop('density', p)
        gz = prism.gz(xp, yp, zp, mesh)
        np.testing.assert_allclose(y, gz, rtol=1e-9, atol=1e-12 * np.abs(gz).max())
        again = block.gram(w, 1.0)
        assert np.array_equal(again, block.gram(w, 1.0))        # reproducible


def test_gram_rejects_bad_shapes(gpu_lib):
    from fatiando_b200 import prism, mesher, gridder, normal
    mesh = mesher.PrismMesh((0, 1000, 0, 1000, 0, 500), (2, 3, 4))
    xp, yp, zp = gridder.scatter((0, 1000, 0, 1000), 9, z=-10., seed=0)
    with prism.Model(mesh, None, keep_all=True) as model, normal.SensitivityBlock(model, xp, yp, zp) as block:
        with pytest.raises(ValueError):
            block.dot(np.ones(5))
        with pytest.raises(ValueError):
            block.tdot(np.ones(4))
        with pytest.raises(ValueError):
            block.gram(np.ones(3))


def test_commensurate_grid_jacobian_is_bit_identical(gpu_lib, oracle):
    """Opt-in reuse of kernel values on a level grid commensurate with the mesh
    (csrc/prism_grid_reuse.cuh): the same bits as the default node-sharing Jacobian, for gravity
    and magnetic fields, host and device destinations; grids that do not qualify fall back."""
    from fatiando_b200 import prism, mesher, gridder, normal, utils
    mesh = mesher.PrismMesh((0, 4000, 0, 3000, 0, 1500), (5, 6, 8))         # 500 x 500 x 300 m cells
    xp, yp, zp = gridder.regular((-1000, 5000, -500, 3500), (61, 41), z=-150)      # 100 m: 5 steps per cell
    fv = utils.dircos(30, -15)
    with prism.Model(mesh, None, keep_all=True) as model:
        plan = prism.grid_plan(model.nodes, xp, yp, zp)
        assert plan is not None and (plan['kx'], plan['ky'], plan['ni'], plan['nj']) == (5, 5, 61, 41)
        for f in ['gz', 'gxy', 'gzz', 'tf']:
            pm = list(fv) if f == 'tf' else None
            ref = model.jacobian(xp, yp, zp, f, pmag=pm, fvec=fv)
            assert model.last_jacobian_path == 'default'
            got = model.jacobian(xp, yp, zp, f, pmag=pm, fvec=fv, grid_reuse=True)
            assert model.last_jacobian_path == 'grid'
            assert np.array_equal(got, ref), f
        with normal.SensitivityBlock(model, xp, yp, zp, 'gz', grid_reuse=True) as block:
            assert block.path == 'grid'
            assert np.array_equal(block.to_host(), model.jacobian(xp, yp, zp, 'gz'))
        # a grid whose spacing does not divide the cells, scattered points, a tilted level
        for pts in (gridder.regular((-1000, 5000, -500, 3500), (60, 41), z=-150),
                    gridder.scatter((-1000, 5000, -500, 3500), 200, z=-150, seed=1)):
            model.jacobian(*pts, field='gz', grid_reuse=True)
            assert model.last_jacobian_path == 'default'
        zt = zp.copy()
        zt[5] -= 1.0
        assert prism.grid_plan(model.nodes, xp, yp, zt) is None
    # coordinates whose differences are NOT reproduced by translation (0.1 m is not a double):
    # the exactness check refuses, the default path runs
    mesh2 = mesher.PrismMesh((0.1, 4000.1, 0.3, 3000.3, 0, 1500), (5, 6, 8))
    x2, y2, z2 = gridder.regular((-999.9, 5000.1, -499.7, 3500.3), (61, 41), z=-150)
    with prism.Model(mesh2, None, keep_all=True) as model:
        j = model.jacobian(x2, y2, z2, 'gz', grid_reuse=True)
        assert model.last_jacobian_path in ('default', 'grid')
        assert np.array_equal(j, model.jacobian(x2, y2, z2, 'gz'))


def test_commensurate_grid_jacobian_on_the_c4_mesh(gpu_lib):
    """The multi-chunk TMA path (800 cell-row segments of 400 bytes per matrix row) on the C4 mesh:
    bit-identical to the default node-sharing Jacobian, host destination."""
    import bench
    from fatiando_b200 import prism, gridder
    w = bench.make_workload("c4", 1)
    xp, yp, zp = gridder.regular((0, 10000, 0, 5000), (51, 26), z=-100)         # 200 m: one step per cell
    with prism.Model(w['model'], None, keep_all=True) as model:
        ref = model.jacobian(xp, yp, zp, 'gz')
        got = model.jacobian(xp, yp, zp, 'gz', grid_reuse=True)
        assert model.last_jacobian_path == 'grid'
        assert got.shape == (51 * 26, 40000) and np.array_equal(got, ref)
