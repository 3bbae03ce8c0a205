#!/usr/bin/env python
"""
Benchmark of the prism forward model on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5]
    python bench.py --impl reference ...          # the reference's CPU path, same metric

Metric (BASELINE.json): prism x point evaluations per second, fp64.  One "step" is
one pass of the hot path over the whole workload.  Default workload = config C2
(configs[1]): the six gravity-gradient components of a 50x50x20 PrismMesh with
variable density on a 200x200 grid (2e9 pairs/step/GPU).  With N GPUs the
observation grid is N times larger and observation-sharded (weak scaling, no
data-path collective); value = all pairs of all ranks / max-over-ranks device time.

  value    : inputs resident in HBM (device pointers through the C ABI), device time
             from CUDA events on the library's own stream, summed over the K steps.
  e2e      : the public Python API (prism.fields) with HOST numpy arrays: mesh
             flattening, model upload, point H2D, kernels, result D2H, every step.
  evaluation : regular PrismMesh models (C2, C3) are evaluated once per mesh NODE
             (csrc/prism_mesh.cuh); --per-cell forces the general per-cell kernels.
  roofline : FP64 pipe.  achieved = pairs/s x F, F = SURVEY 8(d)'s fused-minimal FLOP
             per pair of the REFERENCE formulation (what the answer costs with
             libdevice-grade log/atan2 per corner); peak = DFMA-chain rate measured
             in this run (MEASURED_PEAKS.json has no FP64 figure).  The kernel uses
             an algebraically collapsed formulation with fewer instructions, so
             achieved/peak can exceed 1; "executed_frac" is the real FP64-pipe
             utilisation = executed FP64 instr/pair (ncu, profiles/) x pairs/s / issue rate.
  cpu_baseline : the reference's own Cython kernel (oracle/_ref, "reference") or the C
             port ("port") on this box's host cores, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "prism_point_evals_per_s"
UNIT = "pairs/s"

# algorithmic work per prism x point pair of the reference
# formulation fused into one pass (FLOP, DFMA = 2) and FP64-pipe slots.
WORK = {
    "c2": dict(flop=2802.0, slots=1922.0, what="6 tensor components"),
    "c2a": dict(flop=2802.0, slots=1922.0, what="6 tensor components"),
    "c3": dict(flop=2820.0, slots=1940.0, what="tf+bx+by+bz"),
    "c4": dict(flop=1478.0, slots=1022.0, what="Jacobian gz (+8 B stored per entry)"),
    "c5": dict(flop=1478.0, slots=1022.0, what="gz"),
    "c5t": dict(flop=2850.0, slots=1970.0, what="gz + 6 tensor components"),
    "c4full": dict(flop=1478.0, slots=1022.0, what="Jacobian gz (+8 B stored per entry)"),
    "c5full": dict(flop=1478.0, slots=1022.0, what="gz"),
    "c5tfull": dict(flop=2850.0, slots=1970.0, what="gz + 6 tensor components"),
}
# FP64-pipe instructions the fast kernels actually execute per pair (ncu
# smsp__inst_executed_pipe_fp64 / pairs, see profiles/); None until measured.
# c4: the Jacobian kernel runs the same pair evaluator as c5; the profiled first row block
# executes 373 because 2/3 of its rows lie on mesh planes (careful variant), see profiles/.
# DRAM traffic of the dominant kernel per launch (ncu dram__bytes_read.sum +
# dram__bytes_write.sum, profiles/r01_*): the model and the points are read once, the
# outputs (<= 2 MB) stay in L2 until the D2H copy.  Bytes per launch at the profiled size.
DRAM_TRAFFIC_PER_LAUNCH = {"c2": 1.47e6, "c2a": 1.47e6, "c3": None, "c4": None, "c5": None, "c5t": None}
EXECUTED_FP64_PER_PAIR = {"c2": 421.7, "c2a": 421.7, "c3": 441.8, "c4": 318.5, "c5": 318.5, "c5t": 523.3,
                          "c5full": 318.5, "c5tfull": 523.3, "c4full": 318.5}   # profiles/r01_*
# the same for the node-sharing kernel of regular meshes, per CELL x point pair
# (profiles/r01_c2_tensor_nodes.txt, r01_c3_mag4_nodes.txt)
EXECUTED_FP64_PER_PAIR_NODES = {"c2": 187.3, "c2a": 187.3, "c3": 205.8,
                                "c4": 109.5, "c4full": 109.5}      # c4: profiles/r01_c4_jacobian_nodes.txt
# and for the surface form of a relief (profiles/r01_c5_gz_surface.txt, r01_c5t_gz_tensor_surface.txt)
EXECUTED_FP64_PER_PAIR_SURFACE = {"c5": 202.5, "c5t": 342.6, "c5full": 202.5, "c5tfull": 342.6}


# --------------------------------------------------------------------------
# workloads (SURVEY.md 8(d) "Synthetic inputs"), all seeded
# --------------------------------------------------------------------------
def make_workload(name, world=1):
    from fatiando_b200 import mesher, gridder, utils
    from tools.synthetic import hill
    rs = np.random.RandomState
    if name == "c2":
        mesh = mesher.PrismMesh((0, 5000, 0, 5000, 0, 2000), (20, 50, 50))
        mesh.addprop('density', rs(0).uniform(-1000, 1000, mesh.size))
        xp, yp, zp = gridder.regular((0, 5000, 0, 5000), (200 * world, 200), z=-150)
        return dict(model=mesh, prop='density', names=['gxx', 'gxy', 'gxz', 'gyy', 'gyz', 'gzz'],
                    xp=xp, yp=yp, zp=zp, fvec=None, kind='forward',
                    label="C2: gxx..gzz of PrismMesh 50x50x20 (5e4 prisms, variable density) on a "
                          "%dx200 grid at z=-150" % (200 * world))
    if name == "c2a":
        # C2 with the grid ON the mesh planes (spacing 25 m, cells 100 m): every 4th
        # grid line coincides with cell faces -- the common "aligned" user setup
        w = make_workload("c2", world)
        w['xp'], w['yp'], w['zp'] = gridder.regular((0, 5000, 0, 5000), (201 * world, 201), z=-150)
        w['label'] = "C2 aligned: gxx..gzz of PrismMesh 50x50x20 on a %dx201 grid whose lines hit cell faces" % (201 * world)
        return w
    if name == "c3":
        mesh = mesher.PrismMesh((0, 10000, 0, 10000, 0, 2000), (10, 100, 100))
        inc, dec = 30, -15
        r = rs(1)
        mi = r.uniform(0.5, 2, mesh.size)
        ri = r.uniform(0, 1, mesh.size)
        ii = r.uniform(-90, 90, mesh.size)
        di = r.uniform(-180, 180, mesh.size)
        d2r = np.pi / 180.
        rem = np.stack([ri * np.cos(d2r * ii) * np.cos(d2r * di),
                        ri * np.cos(d2r * ii) * np.sin(d2r * di), ri * np.sin(d2r * ii)], axis=1)
        mag = utils.ang2vec(mi, inc, dec) + rem
        mesh.addprop('magnetization', mag)
        xp, yp, zp = gridder.regular((0, 10000, 0, 10000), (500 * world, 500), z=-300)
        return dict(model=mesh, prop='magnetization', names=['tf', 'bx', 'by', 'bz'],
                    xp=xp, yp=yp, zp=zp, fvec=utils.dircos(inc, dec), inc=inc, dec=dec, kind='forward',
                    label="C3: tf+bx+by+bz of PrismMesh 100x100x10 (1e5 prisms, induced+remanent) "
                          "on a %dx500 grid at z=-300" % (500 * world))
    if name == "c4":
        mesh = mesher.PrismMesh((0, 10000, 0, 10000, 0, 4000), (16, 50, 50))
        # one GPU holds a row block of J: 31250 x 40000 fp64 = 10 GB (1/8 of the 80 GB matrix)
        xp, yp, zp = gridder.regular((0, 10000, 0, 10000), (500, 500), z=-100)
        rows = 31250
        idx = np.arange(rows * world)
        return dict(model=mesh, prop=None, names=['gz'], xp=xp[idx], yp=yp[idx], zp=zp[idx],
                    fvec=None, kind='jacobian',
                    label="C4: gz sensitivity matrix, %d rows (of 250000) x 40000 PrismMesh cells, "
                          "device-resident" % (rows * world))
    if name == "c4full":
        # the whole 250000 x 40000 sensitivity matrix (80 GB) resident on ONE B200, or its
        # row shards on several
        w = make_workload("c4", 1)
        xp, yp, zp = gridder.regular((0, 10000, 0, 10000), (500, 500), z=-100)
        w['xp'], w['yp'], w['zp'] = xp, yp, zp
        w['label'] = "C4 full: gz sensitivity matrix, 250000 observations x 40000 PrismMesh cells (80 GB fp64), device-resident, row-sharded over the GPUs in use"
        w['scaling'] = 'strong'
        return w
    if name in ("c5full", "c5tfull"):
        # the north-star problem itself: ALL 1e6 points x 1e6 prisms, points sharded over
        # the GPUs in use (strong scaling: the total work does not grow with N)
        w = make_workload("c5" if name == "c5full" else "c5t", 1)
        area = (0, 1e6, 0, 1e6)
        w['xp'], w['yp'], w['zp'] = gridder.regular(area, (1000, 1000), z=-5000)
        w['label'] = "C5 full: %s of a 1000x1000 PrismRelief (1e6 prisms) at all 1e6 points, z=-5000" % '+'.join(w['names'])
        w['scaling'] = 'strong'
        return w
    if name in ("c5", "c5t"):
        area, shape = (0, 1e6, 0, 1e6), (1000, 1000)
        x, y = gridder.regular(area, shape)
        r = rs(2)
        h = np.zeros(x.size)
        for _ in range(8):
            h += r.uniform(500, 4000) * hill(x, y, r.uniform(5e4, 2e5), r.uniform(5e4, 2e5),
                                             (r.uniform(0, 1e6), r.uniform(0, 1e6)), r.uniform(0, 180))
        h += r.uniform(0, 200, x.size)
        relief = mesher.PrismRelief(0, gridder.spacing(area, shape), (x, y, -h))
        relief.props['density'] = np.where(-h > 0, -2670.0, 2670.0)   # addprop's sign rule, vectorised
        # 1/64 of the 1e6 points per GPU per step keeps a default run within minutes
        xp, yp, zp = gridder.regular(area, (1000, 1000), z=-5000)
        take = 15625 * world
        idx = np.linspace(0, xp.size - 1, take).astype(np.int64)
        names = ['gz'] if name == "c5" else ['gz', 'gxx', 'gxy', 'gxz', 'gyy', 'gyz', 'gzz']
        return dict(model=relief, prop='density', names=names, xp=xp[idx].copy(), yp=yp[idx].copy(),
                    zp=zp[idx].copy(), fvec=None, kind='forward',
                    label="C5: %s of a 1000x1000 PrismRelief (1e6 prisms) at %d of the 1e6 points, "
                          "z=-5000" % ('+'.join(names), take))
    raise SystemExit("unknown workload %s" % name)


# --------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------
class ClockSampler(object):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "25"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            # nvidia-smi needs a moment before its first line: wait for it, so that even a
            # timed region of a few tens of milliseconds is sampled
            t0 = time.perf_counter()
            while not self.rows and time.perf_counter() - t0 < 2.0:
                time.sleep(0.01)
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Start of the timed region: keep at most the last warm-up sample before it."""
        self.rows = self.rows[-1:]

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------
# CPU baseline: the reference's own kernel on this box's host cores
# --------------------------------------------------------------------------
def _cpu_chunk(args):
    names, xp, yp, zp, bounds, props, fvec, kind = args
    from oracle import prism_oracle as po
    use_ref = po.ref_available()
    out = 0.0
    for f in names:
        pr = props
        if f == 'tf':
            pr = np.hstack([props, np.tile(fvec, (props.shape[0], 1))])
        if kind == 'jacobian':
            # one reference call per cell (eqlayer.py:108-111): a column each
            unit = np.ones((1, 1))
            for q in range(bounds[0].size):
                bq = [b[q:q + 1] for b in bounds]
                col = po.ref_model(f, xp, yp, zp, bq, unit) if use_ref else po.model(f, xp, yp, zp, bq, unit)
                out += float(col[0])
        else:
            r = po.ref_model(f, xp, yp, zp, bounds, pr) if use_ref else po.model(f, xp, yp, zp, bounds, pr)
            out += float(r[0])
    return out


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_baseline(w, flat, seconds=15.0, cores=None):
    """pairs/s of the reference CPU path on a bounded sample of workload w."""
    import multiprocessing as mp
    from oracle import prism_oracle as po
    cores = cores or os.cpu_count() or 1
    m = flat.size
    if w['prop'] == 'magnetization':
        props = flat.magnetization
    elif flat.density is not None:
        props = flat.density[:, None]
    else:
        props = np.ones((m, 1))
    bounds = flat.bounds
    per_field_rate = 1.5e6       # pairs/s/core per field call, sizing only
    pts_per_core = max(1, int(per_field_rate * seconds / (m * len(w['names']))))
    n_s = min(w['xp'].size, pts_per_core * cores)
    idx = np.linspace(0, w['xp'].size - 1, n_s).astype(np.int64)
    xs, ys, zs = w['xp'][idx].copy(), w['yp'][idx].copy(), w['zp'][idx].copy()
    edges = np.linspace(0, n_s, min(cores, n_s) + 1).astype(int)
    jobs = [(w['names'], xs[a:b], ys[a:b], zs[a:b], bounds, props, w['fvec'], w['kind'])
            for a, b in zip(edges[:-1], edges[1:]) if b > a]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(jobs)) as pool:
        warm = [(j[0], j[1][:1], j[2][:1], j[3][:1]) + j[4:] for j in jobs]
        pool.map(_cpu_chunk, warm)                  # import + page-in, one point per worker
        t0 = time.perf_counter()
        pool.map(_cpu_chunk, jobs)
        dt = time.perf_counter() - t0
        # (i) of BASELINE.md section 3: one core, on one worker's share of the sample
        t1 = time.perf_counter()
        pool.map(_cpu_chunk, jobs[:1])
        dt1 = time.perf_counter() - t1
    pairs = float(n_s) * m
    one_core = float(jobs[0][1].size) * m / dt1
    return {"value": pairs / dt, "unit": UNIT, "cores": len(jobs),
            "kind": "reference" if po.ref_available() else "port",
            "cpu_model": _cpu_model(), "one_core_value": one_core,
            "sample": "%d of %d points x all %d prisms, %d field call(s) each (%s), %.1f s on %d cores; "
                      "one core: %d points, %.1f s" % (
                          n_s, w['xp'].size, m, len(w['names']), '+'.join(w['names']), dt, len(jobs),
                          jobs[0][1].size, dt1),
            "seconds": dt}


# --------------------------------------------------------------------------
def dist_setup(gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


def run_reference(args):
    world, rank, _ = dist_setup(args.gpus)
    if rank != 0:
        return
    from fatiando_b200.flatten import flatten
    w = make_workload(args.workload, 1)
    flat = flatten(w['model'], w['prop'], override=w['prop'] is None, fvec=w['fvec'])
    steps = []
    for i in range(args.warmup + args.steps):
        r = cpu_baseline(w, flat, seconds=max(2.0, 60.0 / (args.warmup + args.steps)))
        if i >= args.warmup:
            steps.append(r)
    value = float(np.mean([s['value'] for s in steps]))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * float(np.mean([s['seconds'] for s in steps])),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (seeded)", "config": {"workload": w['label']},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": steps[0]['cores'],
                             "kind": steps[0]['kind'], "sample": steps[0]['sample'],
                             "cpu_model": steps[0]['cpu_model'],
                             "one_core_value": float(np.mean([s['one_core_value'] for s in steps]))},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import torch
    from fatiando_b200 import _lib, prism
    from fatiando_b200.flatten import flatten
    import ctypes

    world, rank, local = dist_setup(args.gpus)
    prism.NODE_SHARING = not args.per_cell          # the e2e leg goes through prism.fields
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.require_device()
    dev = local
    lib = _lib.lib()

    w = make_workload(args.workload, world)
    from fatiando_b200.partition import local_slice
    sl = local_slice(w['xp'].size, rank, world)
    xp, yp, zp = (np.ascontiguousarray(a[sl]) for a in (w['xp'], w['yp'], w['zp']))
    n = xp.size
    flat = flatten(w['model'], w['prop'], override=w['prop'] is None, fvec=w['fvec'])
    m = flat.size
    names = w['names']
    ncomp = len(names)
    # hand the mesh object itself over, so that a regular PrismMesh also gets its
    # node-sharing form (the per-cell arrays are uploaded either way)
    model = prism.Model(w['model'], w['prop'], keep_all=w['prop'] is None, fvec=w['fvec'], device=dev,
                        node_sharing=not args.per_cell)
    tdev = torch.device("cuda", dev)
    d_pts = [torch.from_numpy(a).to(tdev) for a in (xp, yp, zp)]
    ids = sorted(_lib.FIELD_ID[nm] for nm in names)
    mask = sum(1 << i for i in ids)
    scale = np.array([prism.unit_scale(_lib.FIELDS[i]) for i in ids])
    fvec = None if w['fvec'] is None else np.ascontiguousarray(w['fvec'], dtype=np.float64)
    if w['kind'] == 'jacobian':
        d_out = torch.empty((n, m), dtype=torch.float64, device=tdev)
    else:
        d_out = torch.empty((ncomp, n), dtype=torch.float64, device=tdev)
    unit = np.array([1.0])
    torch.cuda.synchronize()

    def step_device():
        if w['kind'] == 'jacobian':
            rc = lib.fb200_jacobian(model._handle, d_pts[0].data_ptr(), d_pts[1].data_ptr(),
                                    d_pts[2].data_ptr(), n, ids[0], _lib.dptr(unit), _lib.dptr(fvec),
                                    float(scale[0]), d_out.data_ptr(), m, _lib.DEVICE_POINTERS)
        else:
            rc = lib.fb200_forward(model._handle, d_pts[0].data_ptr(), d_pts[1].data_ptr(),
                                   d_pts[2].data_ptr(), n, mask, None, None, _lib.dptr(fvec),
                                   _lib.dptr(scale), d_out.data_ptr(), n, _lib.DEVICE_POINTERS)
        _lib.check(rc)
        return model.last_kernel_ms()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the clock sampler needs ~0.1 s to produce its first line: start it before the warm-up
    # (same load) so that short timed regions are covered too
    sampler = ClockSampler(dev)
    sampler.start()
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler.mark()
    dev_ms, launches = 0.0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _lib.check(lib.fb200_flush_l2(dev))        # L2 hygiene, outside the step's events
        ms, nl = step_device()
        dev_ms += ms
        launches += nl
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()

    # ---- e2e: public API, host buffers, everything inside the timed region ----
    def step_e2e():
        if w['kind'] == 'jacobian':
            rows = min(n, 2048)      # host-resident J is PCIe-bound: bounded row block
            j = prism.jacobian(xp[:rows], yp[:rows], zp[:rows], w['model'], names[0], device=dev)
            return rows, j.nbytes
        res = prism.fields(xp, yp, zp, w['model'], names, inc=w.get('inc'), dec=w.get('dec'),
                           device=dev)
        return n, sum(v.nbytes for v in res.values())
    if args.no_e2e:
        rows_done, d2h, e2e_s = n, 0, float('inf')
    else:
        step_e2e()
        barrier()
        t1 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(e2e_steps):
            rows_done, d2h = step_e2e()
        barrier()
        e2e_s = (time.perf_counter() - t1) / e2e_steps
    h2d = (6 + (3 if w['prop'] == 'magnetization' else (1 if w['prop'] else 0))) * m * 8 + 3 * rows_done * 8
    # the merged forms are uploaded with the model every e2e step as well
    if model.nodes is not None:
        h2d += sum(a.nbytes for a in model.nodes.weights) + (model.nodes.xn.nbytes + model.nodes.yn.nbytes
                                                             + model.nodes.zn.nbytes)
    if model.surface is not None:
        h2d += sum(a.nbytes for a in model.surface.faces) + sum(a.nbytes for a in model.surface.nodes)

    # ---- reduce over ranks: max time, summed work ----
    stats = torch.tensor([dev_ms, e2e_s, wall], dtype=torch.float64, device=tdev)
    work = torch.tensor([float(n) * m, float(rows_done) * m, float(launches)], dtype=torch.float64, device=tdev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_max, wall_max = stats.tolist()
    pairs_step, pairs_e2e, launches_all = work.tolist()
    ms_per_step = dev_ms_max / args.steps
    value = pairs_step / (ms_per_step * 1e-3)

    line = None
    if rank == 0:
        wk = WORK[args.workload]
        burst, sustained = ctypes.c_double(0), ctypes.c_double(0)
        _lib.check(lib.fb200_fp64_peak(dev, 1.0, ctypes.byref(burst), ctypes.byref(sustained)))
        per_gpu = value / world
        achieved = per_gpu * wk['flop'] / 1e12
        peak = sustained.value if ms_per_step > 500 else burst.value
        executed = (EXECUTED_FP64_PER_PAIR_NODES if model.nodes is not None
                    else EXECUTED_FP64_PER_PAIR_SURFACE if model.surface is not None and w['kind'] == 'forward'
                    else EXECUTED_FP64_PER_PAIR).get(args.workload)
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak if peak else None,
                    "traffic": DRAM_TRAFFIC_PER_LAUNCH.get(args.workload),
                    "peak_source": "DFMA-chain microbenchmark in this run (fb200_fp64_peak: burst %.2f, "
                                   "sustained %.2f TFLOP/s); MEASURED_PEAKS.json has no FP64 figure"
                                   % (burst.value, sustained.value),
                    "algorithmic_flop_per_pair": wk['flop'], "algorithmic_of": wk['what'],
                    "note": "achieved = pairs/s x F with F = SURVEY 8(d)'s FLOP per pair of the reference "
                            "formulation (libm-grade log/atan2 per corner); the kernels reach the same "
                            "answer with collapsed sums / node sharing in ~190-520 FP64 instructions per "
                            "pair, so frac exceeds 1; executed_frac = executed FP64-pipe instructions "
                            "(ncu, profiles/) x pairs/s / (148 SM x 64 lanes x clock) is the measured "
                            "FP64-pipe utilisation",
                    "executed_fp64_inst_per_pair": executed,
                    "executed_frac": (per_gpu * executed / (148 * 64 * sm_hz)
                                      if executed else None)}
        if w['kind'] == 'jacobian':
            peaks = {}
            try:
                peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            except Exception:
                pass
            hbm = peaks.get("hbm_gbs", 6650.0)
            roofline["hbm_write"] = {"achieved_gbs": per_gpu * 8 / 1e9, "peak_gbs": hbm,
                                     "frac": per_gpu * 8 / 1e9 / hbm,
                                     "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}
        cpu = None
        if world == 1 and not args.no_cpu:
            w1 = make_workload(args.workload, 1)
            cpu = cpu_baseline(w1, flat, seconds=15.0)
            cpu.pop("seconds", None)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": w.get('scaling', 'weak'), "vs_baseline": None,
                "dtype": "f64", "data": "synthetic (seeded numpy RandomState)",
                "config": {"workload": w['label'], "pairs_per_step": pairs_step,
                           "fields": names, "parallelism": "obs-shard x%d" % world,
                           "evaluation": ("node-sharing (regular mesh: one kernel evaluation per node)"
                                          if model.nodes is not None and w['kind'] == 'forward' else
                                          "node-sharing Jacobian (one kernel evaluation per point and node, "
                                          "entries = signed sums of eight node values)"
                                          if model.nodes is not None else
                                          "surface form (relief: top faces + merged reference-plane nodes)"
                                          if model.surface is not None and w['kind'] == 'forward' else
                                          "per cell"),
                           "l2": "flushed between steps (256 MiB memset); inputs (<4 MB) are "
                                 "smem/L2-resident by design, kernel is FP64-bound"},
                "wall_ms_per_step": 1e3 * wall_max / args.steps,
                "e2e": {"value": pairs_e2e / e2e_max, "unit": UNIT,
                        "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "api": "fatiando_b200.prism (flatten + upload + H2D + kernels + D2H per step)"},
                "gpu_launches": int(launches_all),
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    model.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORK))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer e2e leg (huge workloads)")
    ap.add_argument("--per-cell", action="store_true",
                    help="evaluate regular meshes cell by cell (no node sharing)")
    args = ap.parse_args()
    if args.impl == "b200" and not args.workload.endswith("full"):
        args.warmup = max(args.warmup, 3)       # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
