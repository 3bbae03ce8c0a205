"""
Sharding of the prism forward model over several GPUs.

The reference's prism path is serial; its only parallel code splits the
OBSERVATION POINTS into equal contiguous chunks and stacks the results
(tesseroid.py:189-204, chunking :245-266).  Points are independent
(_prism.pyx:265 has no cross-point dependency), so that is the natural unit:

* **observation sharding** (default): every GPU holds the whole flattened model
  (56-80 bytes per prism) and evaluates a contiguous slice of the points.  No
  collective on the data path; results are bit-identical to a single-GPU run.
  For the Jacobian this is row sharding of ``J``.
* **prism sharding** (few points, huge model): every GPU evaluates a contiguous
  slice of the prisms at all points; the partial fields are summed with ONE
  all-reduce (NCCL over NVLink when the tensors live on the GPUs).  This
  changes the summation order, so it equals the single-GPU result only to
  within the parity floor; ranks are summed in rank order by the
  collective's fixed reduction tree, so reruns are reproducible.

Two drivers: one process per GPU under ``torch.distributed`` (the launch the
benchmark uses), or one host thread per GPU inside a single process
(:func:`fields_multi_device`; ctypes releases the GIL during the kernels).

The ``evaluate`` hooks exist so that the host-side logic (slicing, gathering,
reducing) can be tested with ``gloo`` on machines without a GPU; the product
default is always the CUDA path.
"""
import threading

import numpy as np

from . import prism as _prism_api
from .flatten import FlatModel, flatten


def chunk_bounds(n, parts):
    """
    ``parts + 1`` offsets of equal contiguous chunks of ``range(n)``; the last
    chunk takes the remainder (like tesseroid.py:260-265).
    """
    parts = max(1, int(parts))
    size = n // parts
    bounds = [i * size for i in range(parts)] + [n]
    return bounds


def local_slice(n, rank, world):
    """The slice of ``range(n)`` owned by ``rank`` of ``world``."""
    b = chunk_bounds(n, world)
    return slice(b[rank], b[rank + 1])


def default_device(rank=0):
    """
    The CUDA device of this process: LOCAL_RANK under torchrun (the global rank is not a
    device index on a multi-node job), else the rank modulo the visible devices.
    """
    import os
    if 'LOCAL_RANK' in os.environ:
        return int(os.environ['LOCAL_RANK'])
    from . import _lib
    count = _lib.device_count()
    return rank % count if count else 0


def _dist():
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    return dist


def _gpu_evaluate(device):
    def evaluate(flat, prop, xp, yp, zp, names, **kw):
        if prop == 'magnetization' and kw.get('pmag') is None and flat.magnetization is None:
            raise ValueError("model has no magnetization")
        with _prism_api.Model(flat, prop, device=device) as model:
            return model.fields(xp, yp, zp, names, **kw)
    return evaluate


def _flatten_for(prisms, names, dens, pmag, fvec):
    magnetic = any(nm in _prism_api.MAGNETIC_FIELDS for nm in names)
    gravity = any(nm in _prism_api.GRAVITY_FIELDS for nm in names)
    if magnetic and gravity:
        raise ValueError("shard gravity and magnetic fields in separate calls")
    prop = 'magnetization' if magnetic else 'density'
    override = (pmag if magnetic else dens) is not None
    if isinstance(prisms, FlatModel):
        return prisms, prop
    return flatten(prisms, prop, override=override, fvec=fvec), prop


def fields_observation_sharded(xp, yp, zp, prisms, names, dens=None, pmag=None, fvec=None,
                               gather=True, device=None, evaluate=None, **kw):
    """
    Evaluate ``names`` with the points split over the ranks of the default
    ``torch.distributed`` group.  No data-path collective; with ``gather=True``
    the slices are all-gathered so every rank returns the full arrays (the
    ``numpy.hstack`` of the reference's pattern), otherwise each rank returns
    its own slice (and the slice object).
    """
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    flat, prop = _flatten_for(prisms, names, dens, pmag, fvec)
    sl = local_slice(xp.size, rank, world)
    if evaluate is None:
        evaluate = _gpu_evaluate(default_device(rank) if device is None else device)
    mine = evaluate(flat, prop, np.ascontiguousarray(xp[sl]), np.ascontiguousarray(yp[sl]),
                    np.ascontiguousarray(zp[sl]), names, dens=dens, pmag=pmag, fvec=fvec, **kw)
    if not gather:
        return mine, sl
    parts = [None] * world
    dist.all_gather_object(parts, {k: np.asarray(v) for k, v in mine.items()})
    return {nm: np.hstack([p[nm] for p in parts]) for nm in mine}


def fields_prism_sharded(xp, yp, zp, prisms, names, dens=None, pmag=None, fvec=None,
                         device=None, evaluate=None, **kw):
    """
    Evaluate ``names`` with the PRISMS split over the ranks; the partial fields
    are summed with one all-reduce.  Unit factors are applied after the sum.
    """
    import torch
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    flat, prop = _flatten_for(prisms, names, dens, pmag, fvec)
    sl = local_slice(flat.size, rank, world)
    part = FlatModel([b[sl] for b in flat.bounds],
                     None if flat.density is None else flat.density[sl],
                     None if flat.magnetization is None else flat.magnetization[sl],
                     flat.index[sl], flat.total)
    if evaluate is None:
        evaluate = _gpu_evaluate(default_device(rank) if device is None else device)
    mine = evaluate(part, prop, xp, yp, zp, names, dens=dens, pmag=pmag, fvec=fvec,
                    scaled=False, **kw)
    order = sorted(mine)
    stacked = torch.from_numpy(np.stack([np.asarray(mine[nm]) for nm in order]))
    backend = dist.get_backend()
    if backend == 'nccl':
        dev = torch.device('cuda', torch.cuda.current_device())
        stacked = stacked.to(dev)
    dist.all_reduce(stacked, op=dist.ReduceOp.SUM)
    total = stacked.cpu().numpy()
    return {nm: total[i] * _prism_api.unit_scale(nm) for i, nm in enumerate(order)}


def _gpu_jacobian(device):
    def evaluate(flat, xp, yp, zp, field, pmag=None, fvec=None, **kw):
        with _prism_api.Model(flat, None, device=device) as model:
            return model.jacobian(xp, yp, zp, field, pmag=pmag, fvec=fvec, **kw)
    return evaluate


def _gpu_adjoint(device):
    def evaluate(flat, xp, yp, zp, data, field, pmag=None, fvec=None, **kw):
        with _prism_api.Model(flat, None, device=device) as model:
            return model.adjoint(xp, yp, zp, data, field, pmag=pmag, fvec=fvec, **kw)
    return evaluate


def jacobian_row_sharded(xp, yp, zp, prisms, field='gz', pmag=None, fvec=None, device=None,
                         evaluate=None, **kw):
    """
    The rows of the sensitivity matrix that belong to this rank's slice of the
    observation points: ``(J[sl, :], sl)``.  Row sharding is the multi-GPU layout of
    ``J`` (C4: 80 GB = 10 GB per GPU at 8); nothing is exchanged -- consumers work on
    their row block (``J_r @ p`` is the local slice of the predicted data, ``J_r.T @ d_r``
    needs one all-reduce, see :func:`adjoint_point_sharded`).
    """
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    flat = prisms if isinstance(prisms, FlatModel) else flatten(prisms, None, override=True)
    sl = local_slice(xp.size, rank, world)
    if evaluate is None:
        evaluate = _gpu_jacobian(default_device(rank) if device is None else device)
    rows = evaluate(flat, np.ascontiguousarray(xp[sl]), np.ascontiguousarray(yp[sl]),
                    np.ascontiguousarray(zp[sl]), field, pmag=pmag, fvec=fvec, **kw)
    return rows, sl


def adjoint_point_sharded(xp, yp, zp, prisms, data, field='gz', pmag=None, fvec=None, device=None,
                          evaluate=None, **kw):
    """
    ``J.T @ data`` with the observation points split over the ranks: every rank forms the
    adjoint of its own points on the fly (J is never stored) and ONE all-reduce adds the
    per-cell partial sums (NCCL when the tensors live on the GPUs).  This is the exchange
    step of a row-sharded ``J``; the sum order differs from a single-GPU run, so results
    agree to rounding.
    """
    import torch
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    flat = prisms if isinstance(prisms, FlatModel) else flatten(prisms, None, override=True)
    sl = local_slice(xp.size, rank, world)
    if evaluate is None:
        evaluate = _gpu_adjoint(default_device(rank) if device is None else device)
    mine = evaluate(flat, np.ascontiguousarray(xp[sl]), np.ascontiguousarray(yp[sl]),
                    np.ascontiguousarray(zp[sl]), np.ascontiguousarray(data[sl]), field,
                    pmag=pmag, fvec=fvec, **kw)
    total = torch.from_numpy(np.ascontiguousarray(mine, dtype=np.float64))
    if dist.get_backend() == 'nccl':
        total = total.to(torch.device('cuda', torch.cuda.current_device()))
    dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return total.cpu().numpy()


def fields_multi_device(xp, yp, zp, prisms, names, devices, dens=None, pmag=None, fvec=None,
                        **kw):
    """
    Single-process observation sharding: one host thread per device, each with
    its own resident copy of the model; results concatenated in point order.
    """
    flat, prop = _flatten_for(prisms, names, dens, pmag, fvec)
    devices = list(devices)
    bounds = chunk_bounds(xp.size, len(devices))
    results = [None] * len(devices)
    errors = []

    def work(i, dev):
        try:
            sl = slice(bounds[i], bounds[i + 1])
            results[i] = _gpu_evaluate(dev)(flat, prop, np.ascontiguousarray(xp[sl]),
                                            np.ascontiguousarray(yp[sl]),
                                            np.ascontiguousarray(zp[sl]), names,
                                            dens=dens, pmag=pmag, fvec=fvec, **kw)
        except Exception as exc:   # surfaced below, never swallowed
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(i, d)) for i, d in enumerate(devices)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return {nm: np.hstack([r[nm] for r in results]) for nm in results[0]}
