"""
Sweep the split target (thread blocks per SM the prism / node split aims for) over the bench
workloads in ONE process: device-resident inputs, CUDA-event kernel time from the library,
L2 flushed between steps, like bench.py's `value` leg.

    python tools/split_sweep.py [targets, comma separated] [workloads, comma separated] [steps]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    targets = [int(t) for t in (sys.argv[1] if len(sys.argv) > 1 else "20,32,48,64,96").split(",")]
    names = (sys.argv[2] if len(sys.argv) > 2 else "c2,c3,c5t").split(",")
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    import torch
    from fatiando_b200 import _lib, prism
    lib = _lib.lib()
    dev = torch.device("cuda", 0)
    for name in names:
        w = bench.make_workload(name, 1)
        xp, yp, zp = (np.ascontiguousarray(w[k]) for k in ("xp", "yp", "zp"))
        n = xp.size
        ids = sorted(_lib.FIELD_ID[nm] for nm in w["names"])
        mask = 0
        for i in ids:
            mask |= 1 << i
        scale = np.array([prism.unit_scale(_lib.FIELDS[i]) for i in ids])
        fvec = None if w["fvec"] is None else np.ascontiguousarray(w["fvec"], dtype=np.float64)
        d_pts = [torch.from_numpy(a).to(dev) for a in (xp, yp, zp)]
        d_out = torch.empty((len(ids), n), dtype=torch.float64, device=dev)
        with prism.Model(w["model"], w["prop"], fvec=fvec, device=0) as model:
            ref = None
            for tgt in targets:
                os.environ["FB200_SPLIT_TARGET"] = str(tgt)
                ms = []
                for it in range(steps + 2):
                    _lib.check(lib.fb200_flush_l2(0))
                    _lib.check(lib.fb200_forward(model._handle, d_pts[0].data_ptr(), d_pts[1].data_ptr(),
                                                 d_pts[2].data_ptr(), n, mask, None, None, _lib.dptr(fvec),
                                                 _lib.dptr(scale), d_out.data_ptr(), n, _lib.DEVICE_POINTERS))
                    if it >= 2:
                        ms.append(model.last_kernel_ms()[0])
                torch.cuda.synchronize()
                out = d_out.cpu().numpy()
                if ref is None:
                    ref = out.copy()
                dev_rel = float(np.max(np.abs(out - ref)) / np.max(np.abs(ref)))
                print("%-4s target %4d  %9.3f ms  (min %9.3f)  %.4e pairs/s  max |diff| vs first target %.1e (relative to max |field|)"
                      % (name, tgt, np.mean(ms), np.min(ms), n * model.size / (np.mean(ms) * 1e-3), dev_rel), flush=True)


if __name__ == "__main__":
    main()
