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
    if dist.get_backend() != 'nccl':
        # gloo (CPU tests): small arrays, pickled
        parts = [None] * world
        dist.all_gather_object(parts, {k: np.asarray(v) for k, v in mine.items()})
        return {nm: np.hstack([p[nm] for p in parts]) for nm in mine}
    # NCCL: one all-gather of a float64 tensor padded to the largest chunk (the last chunk takes
    # the remainder), no pickling through the GPU
    import torch
    order = list(mine)
    bounds = chunk_bounds(xp.size, world)
    widest = max(b - a for a, b in zip(bounds[:-1], bounds[1:]))
    dev = torch.device('cuda', default_device(rank) if device is None else device)
    local = torch.zeros((len(order), widest), dtype=torch.float64, device=dev)
    stacked = np.stack([np.asarray(mine[nm], dtype=np.float64) for nm in order])
    local[:, :stacked.shape[1]] = torch.from_numpy(stacked).to(dev)
    everything = torch.empty((world, len(order), widest), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(everything, local)
    host = everything.cpu().numpy()
    return {nm: np.hstack([host[r, i, :bounds[r + 1] - bounds[r]] for r in range(world)])
            for i, nm in enumerate(order)}


class PrismShards(object):
    """
    Prism-sharded evaluation, device-resident: this rank's contiguous slice of the prisms is
    uploaded ONCE; :meth:`fields` evaluates the partial fields of all points straight into a
    GPU tensor (``FB200_DEVICE_POINTERS``), adds the ranks' partials with ONE all-reduce on the
    GPUs (NCCL over NVLink) and applies the unit factors to the sum.  Nothing crosses PCIe
    between the kernel and the collective.  Ranks are added by the collective's fixed tree:
    reruns reproduce, and the result equals a single-GPU run to within the parity floor (the
    summation order differs).

    ``timings`` of the last call: kernel milliseconds (library stream, CUDA events), collective
    milliseconds (torch stream, CUDA events) and the bytes each rank contributed.
    """

    def __init__(self, prisms, names, dens=None, pmag=None, fvec=None, device=None):
        import torch
        dist = _dist()
        self.dist, self.torch = dist, torch
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        flat, prop = _flatten_for(prisms, names, dens, pmag, fvec)
        sl = local_slice(flat.size, self.rank, self.world)
        part = FlatModel([b[sl] for b in flat.bounds],
                         None if flat.density is None else flat.density[sl],
                         None if flat.magnetization is None else flat.magnetization[sl],
                         flat.index[sl], flat.total)
        self.device = default_device(self.rank) if device is None else device
        self.names = list(names)
        self.prop, self.total_prisms, self.slice = prop, flat.size, sl
        self.fvec = None if fvec is None else np.ascontiguousarray(fvec, dtype=np.float64)
        self.model = _prism_api.Model(part, prop, fvec=fvec, device=self.device, node_sharing=False)
        self.timings = {}

    def close(self):
        self.model.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def fields_device(self, d_xp, d_yp, d_zp):
        """``[ncomp, n]`` float64 CUDA tensor (rows in ascending field order, see ``order``) for
        float64 CUDA tensors of point coordinates on this rank's device."""
        from . import _lib
        torch, dist = self.torch, self.dist
        n = d_xp.numel()
        ids = sorted(set(_lib.FIELD_ID[nm] for nm in self.names))
        self.order = [_lib.FIELDS[i] for i in ids]
        mask = 0
        for i in ids:
            mask |= 1 << i
        out = torch.empty((len(ids), n), dtype=torch.float64, device=d_xp.device)
        ones = np.ones(len(ids))
        if self.model.size:
            _lib.check(_lib.lib().fb200_forward(
                self.model._handle, d_xp.data_ptr(), d_yp.data_ptr(), d_zp.data_ptr(), n, mask, None, None,
                _lib.dptr(self.fvec), _lib.dptr(ones), out.data_ptr(), n, _lib.DEVICE_POINTERS))
            kernel_ms = self.model.last_kernel_ms()[0]
        else:
            out.zero_()
            kernel_ms = 0.0
        # fb200_forward returns after its stream has drained: the partials are complete
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
        e1.record()
        scale = torch.tensor([_prism_api.unit_scale(nm) for nm in self.order], dtype=torch.float64,
                             device=out.device)
        out *= scale[:, None]
        torch.cuda.synchronize(out.device)
        self.timings = {"kernel_ms": kernel_ms, "collective_ms": e0.elapsed_time(e1),
                        "collective_bytes": out.numel() * 8}
        return out

    def fields(self, xp, yp, zp):
        torch = self.torch
        dev = torch.device('cuda', self.device)
        pts = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev) for a in (xp, yp, zp)]
        total = self.fields_device(*pts).cpu().numpy()
        return {nm: total[i] for i, nm in enumerate(self.order)}


def fields_prism_sharded(xp, yp, zp, prisms, names, dens=None, pmag=None, fvec=None,
                         device=None, evaluate=None, **kw):
    """
    Evaluate ``names`` with the PRISMS split over the ranks; the partial fields
    are summed with one all-reduce.  Unit factors are applied after the sum.

    On the GPUs (NCCL backend, no ``evaluate`` hook, no overrides) this is
    :class:`PrismShards`: partials stay in HBM, one NCCL all-reduce, nothing staged through
    the host.  The host path below serves the ``gloo`` tests and overrides.
    """
    import torch
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    if evaluate is None and dist.get_backend() == 'nccl' and dens is None and pmag is None and not kw:
        with PrismShards(prisms, names, fvec=fvec, device=device) as shards:
            return shards.fields(xp, yp, zp)
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
        dev = torch.device('cuda', default_device(rank) if device is None else device)
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
        total = total.to(torch.device('cuda', default_device(rank) if device is None else device))
    dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return total.cpu().numpy()


def gram_row_sharded(xp, yp, zp, prisms, field='gz', weights=None, residual=None, factor=1.0,
                     pmag=None, fvec=None, device=None, to_host=True, evaluate=None):
    """
    The normal-equation pieces of a row-sharded sensitivity matrix (the Gauss-Newton step of
    ``fatiando.inversion``: ``Misfit.hessian`` = ``2 J^T W J``, misfit.py:250-256, and
    ``Misfit.gradient`` = ``-2 J^T W r``, misfit.py:280-294): every rank builds the rows of ITS
    observation points in HBM (:class:`fatiando_b200.normal.SensitivityBlock`), forms
    ``factor * J_r^T W_r J_r`` with the fp64 SYRK kernel and -- when ``residual`` is given --
    ``J_r^T W_r r_r``; ONE all-reduce each (NCCL over NVLink, on the GPU tensors) adds the ranks'
    contributions.  This is the one exchange on this path that moves real data: the full
    ``(ncells, ncells)`` fp64 matrix (12.8 GB at config C4).

    Returns ``(G, g, timings)``: numpy arrays (``to_host=True``) or CUDA tensors; ``g`` is None
    without a residual.  ``timings``: build / gram kernel / collective milliseconds, bytes.

    ``evaluate(flat, x, y, z, field, weights, factor, residual) -> (G_local, g_local)`` replaces
    the CUDA path (host arrays, any backend): the hook of the ``gloo`` tests.
    """
    import torch
    from .normal import SensitivityBlock
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    flat = prisms if isinstance(prisms, FlatModel) else flatten(prisms, None, override=True)
    sl = local_slice(xp.size, rank, world)
    if evaluate is not None:
        G_local, g_local = evaluate(flat, np.ascontiguousarray(xp[sl]), np.ascontiguousarray(yp[sl]),
                                    np.ascontiguousarray(zp[sl]), field,
                                    None if weights is None else np.ascontiguousarray(weights[sl]), factor,
                                    None if residual is None else np.ascontiguousarray(residual[sl]))
        G = torch.from_numpy(np.ascontiguousarray(G_local, dtype=np.float64))
        dist.all_reduce(G, op=dist.ReduceOp.SUM)
        g = None
        if g_local is not None:
            g = torch.from_numpy(np.ascontiguousarray(g_local, dtype=np.float64))
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
        return G.numpy(), None if g is None else g.numpy(), {}
    dev_index = default_device(rank) if device is None else device
    dev = torch.device('cuda', dev_index)
    w_local = None if weights is None else np.ascontiguousarray(weights[sl], dtype=np.float64)
    with _prism_api.Model(flat, None, device=dev_index) as model:
        block = SensitivityBlock(model, np.ascontiguousarray(xp[sl]), np.ascontiguousarray(yp[sl]),
                                 np.ascontiguousarray(zp[sl]), field, pmag=pmag, fvec=fvec)
    try:
        G_dev = block.gram_device(w_local, factor)
        timings = {"build_ms": block.build_ms, "gram_ms": block.last_ms}
        G = torch.as_tensor(G_dev, device=dev)             # zero-copy view (__cuda_array_interface__)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_reduce(G, op=dist.ReduceOp.SUM)
        e1.record()
        torch.cuda.synchronize(dev)
        timings.update(collective_ms=e0.elapsed_time(e1), collective_bytes=G.numel() * 8)
        g = None
        if residual is not None:
            g_local = block.tdot(np.ascontiguousarray(residual[sl], dtype=np.float64), w_local, 1.0)
            g = torch.from_numpy(g_local).to(dev)
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
        if to_host:
            G_out = G.cpu().numpy()
            g_out = None if g is None else g.cpu().numpy()
        else:
            G_out, g_out = G.clone(), g
        del G
        G_dev.close()
        return G_out, g_out, timings
    finally:
        block.close()


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
