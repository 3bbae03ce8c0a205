"""
Multi-GPU check of the partitioner under NCCL (run with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_nccl_check.py

Observation sharding must reproduce the single-GPU result bit for bit (list-order
summation); prism sharding (one NCCL all-reduce) to rounding.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from fatiando_b200 import mesher, gridder, partition, prism
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    rs = np.random.RandomState(3)
    mesh = mesher.PrismMesh((0, 5000, 0, 5000, 0, 2000), (6, 20, 20))
    mesh.addprop('density', rs.uniform(-1000, 1000, mesh.size))
    xp, yp, zp = gridder.regular((0, 5000, 0, 5000), (60, 50), z=-150)
    names = ['gz', 'gxx', 'gxy', 'gzz']
    with prism.Model(mesh, 'density', device=local) as model:
        full = model.fields(xp, yp, zp, names, list_order=True)
    obs = partition.fields_observation_sharded(xp, yp, zp, mesh, names, device=local,
                                               list_order=True)
    pri = partition.fields_prism_sharded(xp, yp, zp, mesh, names, device=local)
    ok = True
    why = []

    def expect(cond, what):
        nonlocal ok
        if not cond:
            ok = False
            why.append(what)
    for nm in names:
        expect(np.array_equal(obs[nm], full[nm]), 'observation sharding %s not bit-identical' % nm)
        scale = np.abs(full[nm]).max()
        expect(np.allclose(pri[nm], full[nm], rtol=1e-10, atol=1e-12 * scale),
               'prism sharding %s: max diff %.3g of %.3g' % (nm, np.abs(pri[nm] - full[nm]).max(), scale))
    # row-sharded Jacobian (no exchange) and the adjoint's one all-reduce
    data = np.random.RandomState(4).normal(size=xp.size)
    with prism.Model(mesh, None, keep_all=True, device=local) as model:
        jac = model.jacobian(xp, yp, zp, 'gz')
        adj = model.adjoint(xp, yp, zp, data, 'gz')
    rows, sl = partition.jacobian_row_sharded(xp, yp, zp, mesh, 'gz', device=local)
    expect(np.array_equal(rows, jac[sl]), 'row-sharded Jacobian not bit-identical')
    got = partition.adjoint_point_sharded(xp, yp, zp, mesh, data, 'gz', device=local)
    # a sum over 3000 points of node values that cancel eight-fold into cells: the two summation
    # orders agree to ~1e-11 of the largest entry (the per-cell floor is checked in test_gpu_parity)
    expect(np.allclose(got, adj, rtol=1e-10, atol=1e-10 * np.abs(adj).max()),
           'point-sharded adjoint: max diff %.3g of %.3g' % (np.abs(got - adj).max(), np.abs(adj).max()))
    if why:
        print("rank %d: %s" % (rank, "; ".join(why)), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("dist_nccl_check world=%d: %s" % (world, "OK" if flag.item() == 1 else "MISMATCH"))
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
