"""
Host-side logic of the multi-GPU partitioner with torch.distributed/gloo on CPU,
world_size 2.  The evaluator is injected (the oracle), because this machine has no
GPU; what is tested is the slicing, gathering and reduction -- the same code paths
bench.py and the product use with the CUDA evaluator under NCCL.
"""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_evaluate(flat, prop, xp, yp, zp, names, dens=None, pmag=None, fvec=None,
                     scaled=True, **kw):
    from oracle import prism_oracle as po
    out = {}
    m = flat.size
    for nm in names:
        if nm in ('tf', 'bx', 'by', 'bz'):
            props = flat.magnetization if pmag is None else np.tile(pmag, (m, 1))
            if nm == 'tf':
                props = np.hstack([props, np.tile(fvec, (m, 1))])
        else:
            props = (flat.density if dens is None else np.full(m, float(dens)))[:, None]
        out[nm] = po.model(nm, xp, yp, zp, flat.bounds, props, scale=None if scaled else 1.0)
    return out


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fatiando_b200 import mesher, gridder, partition
        rs = np.random.RandomState(3)
        mesh = mesher.PrismMesh((0, 5000, 0, 5000, 0, 2000), (3, 5, 7))
        mesh.addprop('density', rs.uniform(-1000, 1000, mesh.size))
        mesh.mask = [2, 11]
        xp, yp, zp = gridder.regular((0, 5000, 0, 5000), (9, 7), z=-150)
        names = ['gz', 'gzz']
        obs = partition.fields_observation_sharded(xp, yp, zp, mesh, names,
                                                   evaluate=_oracle_evaluate)
        mine, sl = partition.fields_observation_sharded(xp, yp, zp, mesh, names, gather=False,
                                                        evaluate=_oracle_evaluate)
        pri = partition.fields_prism_sharded(xp, yp, zp, mesh, names, evaluate=_oracle_evaluate)
        def jac_eval(flat, x, y, z, field, pmag=None, fvec=None, **kw):
            from oracle import prism_oracle as po
            return po.jacobian(field, x, y, z, flat.bounds, [1.0])

        def adj_eval(flat, x, y, z, d, field, pmag=None, fvec=None, **kw):
            return jac_eval(flat, x, y, z, field).T @ d
        data = np.random.RandomState(5).normal(size=xp.size)
        rows, slj = partition.jacobian_row_sharded(xp, yp, zp, mesh, 'gz', evaluate=jac_eval)
        adj = partition.adjoint_point_sharded(xp, yp, zp, mesh, data, 'gz', evaluate=adj_eval)

        def gram_eval(flat, x, y, z, field, w, factor, r):
            J = jac_eval(flat, x, y, z, field)
            return factor * (J.T @ (w[:, None] * J)), J.T @ (w * r)
        wts = np.random.RandomState(6).uniform(0.5, 2.0, xp.size)
        G, g, _ = partition.gram_row_sharded(xp, yp, zp, mesh, 'gz', weights=wts, residual=data, factor=2.0,
                                             evaluate=gram_eval)
        q.put((rank, {k: v.copy() for k, v in obs.items()}, (sl.start, sl.stop),
               {k: v.copy() for k, v in mine.items()}, {k: v.copy() for k, v in pri.items()},
               rows.copy(), (slj.start, slj.stop), adj.copy(), G.copy(), g.copy()))
    finally:
        dist.destroy_process_group()


def test_sharding_world_size_2_gloo(oracle):
    import torch.multiprocessing as mp
    from fatiando_b200 import mesher, gridder
    from fatiando_b200.flatten import flatten
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rs = np.random.RandomState(3)
    mesh = mesher.PrismMesh((0, 5000, 0, 5000, 0, 2000), (3, 5, 7))
    mesh.addprop('density', rs.uniform(-1000, 1000, mesh.size))
    mesh.mask = [2, 11]
    xp, yp, zp = gridder.regular((0, 5000, 0, 5000), (9, 7), z=-150)
    flat = flatten(mesh, 'density')
    full = {nm: oracle.model(nm, xp, yp, zp, flat.bounds, flat.density[:, None])
            for nm in ['gz', 'gzz']}
    covered = np.zeros(xp.size, dtype=int)
    flat_all = flatten(mesh, None, override=True)
    jac = oracle.jacobian('gz', xp, yp, zp, flat_all.bounds, [1.0])
    data = np.random.RandomState(5).normal(size=xp.size)
    wts = np.random.RandomState(6).uniform(0.5, 2.0, xp.size)
    for rank, obs, (a, b), mine, pri, rows, (ja, jb), adj, G, g in results:
        # row-sharded normal equations: one all-reduce of the (m, m) matrix and of the gradient
        np.testing.assert_allclose(G, 2.0 * (jac.T @ (wts[:, None] * jac)), rtol=1e-11,
                                   atol=1e-14 * np.abs(jac).max() ** 2)
        np.testing.assert_allclose(g, jac.T @ (wts * data), rtol=1e-11, atol=1e-14 * np.abs(jac).max())
        # row-sharded Jacobian: each rank holds exactly its rows; the adjoint's all-reduce
        # gives every rank the full J^T d
        assert np.array_equal(rows, jac[ja:jb]) and (ja, jb) == (a, b)
        np.testing.assert_allclose(adj, jac.T @ data, rtol=1e-12, atol=1e-15 * np.abs(jac).max())
        covered[a:b] += 1
        for nm in full:
            # observation sharding: bit-identical to the unsharded run, on every rank
            assert np.array_equal(obs[nm], full[nm])
            assert np.array_equal(mine[nm], full[nm][a:b])
            # prism sharding: one all-reduce; equal to rounding (summation order differs)
            np.testing.assert_allclose(pri[nm], full[nm], rtol=1e-11, atol=1e-14)
    assert np.all(covered == 1)          # contiguous chunks tile the points exactly once


def test_chunk_bounds():
    from fatiando_b200.partition import chunk_bounds, local_slice
    assert chunk_bounds(10, 3) == [0, 3, 6, 10]          # last chunk takes the remainder
    assert chunk_bounds(7, 1) == [0, 7]
    assert chunk_bounds(2, 4) == [0, 0, 0, 0, 2]
    assert local_slice(10, 2, 3) == slice(6, 10)
