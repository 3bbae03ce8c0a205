"""
One command per side for the profiles bench.py's roofline depends on.

On the GPU box (under gpurun, one GPU):
    python tools/profile_all.py gpu [keys...]
        one `ncu --set full --clock-control none --import-source on` capture of the dominant
        kernel of each workload -> gpurun_out/r02_<key>.ncu-rep, plus gpurun_out/r02_kernel_hashes.json
        (per kernel, the hash of the snapshot's sources of that kernel: what its capture is valid for).

Here, afterwards:
    python tools/profile_all.py collect [keys...]
        text summary + profiles/r02_kernel_<key>.json for each capture that came back.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# key: (workload, extra profile_case args, ncu kernel regex, pairs per launch, title)
CASES = {
    "c2_nodes": ("c2", [], "mesh_forward", 4e4 * 5e4,
                 "C2 six tensor components, node-sharing kernel mesh_forward_kernel<0x3F0>, 200x200 points x 50x50x20 mesh"),
    "c3_nodes": ("c3", ["--points", "65536"], "mesh_forward", 65536 * 1e5,
                 "C3 tf+bx+by+bz, node-sharing kernel mesh_forward_kernel<0x3C00>, 65536 of the 500x500 points x 100x100x10 mesh"),
    "c4_nodes": ("c4", ["--rows", "8192"], "mesh_jacobian", 8192 * 4e4,
                 "C4 Jacobian gz, node-sharing kernel mesh_jacobian_kernel<0x008>, 8192 rows x 40000 cells, device-resident"),
    "c4_grid": ("c4a", [], "expand_rows", 31626 * 4e4,
                "C4 mesh on a commensurate 251x126 grid: expand_rows_kernel (TMA data movement), 31626 rows x 40000 cells"),
    "c5_surface": ("c5", ["--points", "4096"], "forward_kernel", 4096 * 1e6,
                   "C5 gz, relief surface form forward_kernel<0x008,EvalFace>, 4096 points x 1e6 top faces"),
    "c5t_surface": ("c5t", ["--points", "4096"], "forward_kernel", 4096 * 1e6,
                    "C5 gz+tensor, relief surface form forward_kernel<0x3F8,EvalFace>, 4096 points x 1e6 top faces"),
    "c2_cells": ("c2", ["--per-cell"], "forward_kernel", 4e4 * 5e4,
                 "C2 six tensor components, general per-cell kernel forward_kernel<0x3F0,EvalFast>"),
    "c2_cells_potential": ("c2", ["--per-cell", "--names", "potential"], "forward_kernel", 4e4 * 5e4,
                           "potential alone, per-cell kernel forward_kernel<0x001,EvalFast> (168 registers, 3 blocks/SM), C2 model and grid"),
    "c2_cells_gxyz": ("c2", ["--per-cell", "--names", "gx,gy,gz"], "forward_kernel", 4e4 * 5e4,
                      "gx+gy+gz, per-cell kernel forward_kernel<0x00E,EvalFast> (3 blocks/SM), C2 model and grid"),
    "c2_cells_all10": ("c2", ["--per-cell", "--names", "potential,gx,gy,gz,gxx,gxy,gxz,gyy,gyz,gzz"], "forward_kernel",
                       4e4 * 5e4, "all ten gravity fields, per-cell kernel forward_kernel<0x3FF,EvalFast> (3 blocks/SM), C2 model and grid"),
    "c3_cells": ("c3", ["--per-cell", "--points", "32768"], "forward_kernel", 32768 * 1e5,
                 "C3 tf+bx+by+bz, per-cell kernel forward_kernel<0x3C00,EvalFast>, 32768 points x 1e5 prisms"),
    "c1_cells": ("c1", [], "forward_kernel", 1e4 * 3,
                 "C1 gz of 3 prisms x 100x100 points, per-cell kernel forward_kernel<0x008,EvalFast>"),
}


SIBLINGS = {"sphere_gz": ("sphere", 1e5 * 1.6e5, "spheres: forward_kernel<0x008,EvalSphere>, 1e5 spheres x 400x400 points"),
            "polyprism_gz": ("polyprism 500", 1e4 * 1.6e5, "polygonal prisms: forward_kernel<0x008,EvalPolyEdge>, 500 prisms x 20 "
                                                         "edges x 400x400 points (rows = edges)")}


def _title_pairs(key):
    if key in SIBLINGS:
        _, pairs, title = SIBLINGS[key]
    else:
        _, _, _, pairs, title = CASES[key]
    return pairs, title


def gpu(keys):
    """Capture, summarise ON THE BOX (the reports are tens of MB each; gpurun brings back 64 MiB),
    keep the text + JSON summaries under gpurun_out/."""
    import bench
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    import json
    hashfile = os.path.join(out, "r02_kernel_hashes.json")
    hashes = {k: bench.kernel_hash(k) for k in bench.KERNEL_TU}
    hashes["launch"] = bench.launch_hash()
    json.dump(hashes, open(hashfile, "w"))
    for key in keys:
        rep = os.path.join(out, "r02_" + key)
        if key in SIBLINGS:
            target = [os.path.join(ROOT, "tools", "bench_siblings.py")] + SIBLINGS[key][0].split()
            regex = "forward_kernel"
        else:
            wl, extra, regex, _, _ = CASES[key]
            target = [os.path.join(ROOT, "tools", "profile_case.py"), wl, "1"] + extra
        cmd = ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on",
               "-k", "regex:" + regex, "-c", "1", "-f", "-o", rep, sys.executable] + target
        r = subprocess.run(cmd, capture_output=True, text=True)
        print(key, "rc", r.returncode, (r.stdout + r.stderr).strip().splitlines()[-1:], flush=True)
        pairs, title = _title_pairs(key)
        if os.path.exists(rep + ".ncu-rep"):
            subprocess.call([sys.executable, os.path.join(ROOT, "tools", "summarize_profile.py"), rep + ".ncu-rep",
                             os.path.join(out, "r02_%s.txt" % key), repr(float(pairs)), title, key, hashfile, out])
            os.remove(rep + ".ncu-rep")


def collect(keys):
    """Copy the summaries that came back into profiles/ (tracked)."""
    import shutil
    out = os.path.join(ROOT, "gpurun_out")
    for key in keys:
        for name in ("r02_%s.txt" % key, "r02_kernel_%s.json" % key):
            src = os.path.join(out, name)
            if os.path.exists(src):
                shutil.copy(src, os.path.join(ROOT, "profiles", name))
                print("profiles/" + name)
            else:
                print("missing", src)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else ""
    keys = sys.argv[2:] or list(CASES) + list(SIBLINGS)
    {"gpu": gpu, "collect": collect}[mode](keys)
