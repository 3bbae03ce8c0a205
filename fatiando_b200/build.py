"""
Build libfatiando_b200.so in-tree with nvcc for sm_100a.

    python -m fatiando_b200.build [--force] [--jobs N]

Every .cu under csrc/ is compiled with
    -gencode arch=compute_100a,code=sm_100a -lineinfo
(cross-compiles without a GPU) and linked into
fatiando_b200/libfatiando_b200.so, which travels to the GPU box with the repo
snapshot.  The reference-order translation units (*_exact.cu) are compiled with
-fmad=false so that every product and sum rounds separately, like the
reference's x86-64 -O2 build.
"""
import concurrent.futures
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libfatiando_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "--expt-relaxed-constexpr",
          "-Xcompiler", "-fPIC", "-I", INCLUDE]


def nvcc():
    path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(path):
        raise RuntimeError("nvcc not found; libfatiando_b200.so cannot be built")
    return path


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, jobs=None, verbose=False, extra_flags=(), tag=""):
    """tag/extra_flags build an experimental variant next to the product library."""
    global BUILD, LIB
    if tag:
        BUILD = os.path.join(HERE, "build_" + tag)
        LIB = os.path.join(HERE, "libfatiando_b200_%s.so" % tag)
    os.makedirs(BUILD, exist_ok=True)
    headers = (glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h"))
               + glob.glob(os.path.join(INCLUDE, "*.h")) + [os.path.abspath(__file__)])
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    objs, todo = [], []
    for src in sources:
        obj = os.path.join(BUILD, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            flags = list(COMMON) + list(extra_flags)
            if src.endswith("_exact.cu"):
                flags.append("-fmad=false")
            if verbose:
                flags += ["-Xptxas", "-v"]
            todo.append([nvcc()] + ARCH + flags + ["-c", src, "-o", obj])

    def run(cmd):
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), res.stderr))
        return res.stderr

    if todo:
        with concurrent.futures.ThreadPoolExecutor(jobs or min(8, os.cpu_count() or 1)) as ex:
            for log in ex.map(run, todo):
                if verbose and log:
                    sys.stderr.write(log)
    if force or todo or _stale(LIB, objs):
        run([nvcc()] + ARCH + ["-shared", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    tag = next((a[6:] for a in sys.argv[1:] if a.startswith("--tag=")), "")
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, extra_flags=extra, tag=tag))
