"""
Throughput of the Gauss-Newton products on a device-resident row block of the C4 sensitivity
matrix: J^T J by the fp64 SYRK kernel against the FP64 peak measured in the
same run, J^T r and J p against the HBM bandwidth of MEASURED_PEAKS.json.

    python tools/bench_normal.py [rows]                 # one GPU
    torchrun --nproc-per-node N tools/bench_normal.py   # row-sharded, one all-reduce of G
"""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np      # noqa: E402
import bench            # noqa: E402
from fatiando_b200 import _lib, prism, normal   # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 31250
    w = bench.make_workload("c4full", 1)
    m = 40000
    out = {"workload": "C4 mesh (40000 cells), %d rows per GPU, %d GPU(s)" % (rows, world)}
    if world > 1:
        import torch
        import torch.distributed as dist
        from fatiando_b200 import partition
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        n = rows * world
        xp, yp, zp = (np.ascontiguousarray(w[k][:n]) for k in ('xp', 'yp', 'zp'))
        for rep in range(2):
            t0 = time.perf_counter()
            G, g, tm = partition.gram_row_sharded(xp, yp, zp, w['model'], 'gz', residual=np.ones(n),
                                                  device=local, to_host=False)
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            del G
        flop = 2.0 * rows * m * (m + 128) / 2
        out.update(tm, wall_s=wall, gram_tflops_per_gpu=flop / tm["gram_ms"] / 1e9,
                   collective_algbw_gbs=tm["collective_bytes"] / 1e9 / (tm["collective_ms"] * 1e-3),
                   collective_busbw_gbs=tm["collective_bytes"] / 1e9 / (tm["collective_ms"] * 1e-3) * 2 * (world - 1) / world)
        if rank == 0:
            print(json.dumps(out))
        dist.destroy_process_group()
        return
    burst, sustained = ctypes.c_double(0), ctypes.c_double(0)
    _lib.check(_lib.lib().fb200_fp64_peak(local, 1.0, ctypes.byref(burst), ctypes.byref(sustained)))
    xp, yp, zp = (np.ascontiguousarray(w[k][:rows]) for k in ('xp', 'yp', 'zp'))
    with prism.Model(w['model'], None, keep_all=True, device=local) as model, \
            normal.SensitivityBlock(model, xp, yp, zp, 'gz') as block:
        G = None
        for rep in range(2):
            if G is not None:
                G.close()
            G = block.gram_device()
        G.close()
        flop = 2.0 * rows * m * (m + 128) / 2              # tiles on and above the diagonal
        out.update(build_ms=block.build_ms, gram_ms=block.last_ms, gram_tflops=flop / block.last_ms / 1e9,
                   fp64_peak_tflops=sustained.value if block.last_ms > 500 else burst.value)
        out["gram_frac_of_fp64_peak"] = out["gram_tflops"] / out["fp64_peak_tflops"]
        r = np.ones(rows)
        p = np.ones(m)
        for name, fn in (("tdot", lambda: block.tdot(r)), ("dot", lambda: block.dot(p))):
            fn()
            dt = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                fn()
                dt = min(dt, time.perf_counter() - t0)
            out[name + "_ms_wall"] = 1e3 * dt
            out[name + "_gbs"] = rows * m * 8 / 1e9 / dt
    print(json.dumps(out))


if __name__ == "__main__":
    main()
