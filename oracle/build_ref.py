"""
Build the reference's own prism kernel, unmodified, as a CPU checker.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Compiles
/root/reference/fatiando/gravmag/_prism.pyx where it lies (cython -> C in a
temporary directory -> gcc -O2, the reference's own flags, setup.py:70-85) and
writes ONLY the resulting extension module into oracle/_ref/.  No reference
source is copied into the repository.  The pre-generated _prism.c shipped with
the reference (Cython 0.21) does not compile against CPython 3.12 / numpy 2,
so the .pyx is re-cythonized.

The GPU box has no /root/reference: it uses the prebuilt oracle/_ref/*.so that
travels with the repo snapshot (oracle/_ref/ is git-ignored, not
gpurun-ignored).  Without the reference tree this script is a no-op.
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PYX = "/root/reference/fatiando/gravmag/_prism.pyx"
OUT_DIR = os.path.join(HERE, "_ref")


def ext_path():
    return os.path.join(OUT_DIR, "_prism" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force=False):
    """Return the path of the built extension, or None when unavailable."""
    out = ext_path()
    if not os.path.exists(REF_PYX):
        return out if os.path.exists(out) else None
    if (not force and os.path.exists(out)
            and os.path.getmtime(out) >= os.path.getmtime(REF_PYX)):
        return out
    import numpy
    os.makedirs(OUT_DIR, exist_ok=True)
    gcc = "/usr/bin/gcc" if os.access("/usr/bin/gcc", os.X_OK) else "gcc"
    with tempfile.TemporaryDirectory() as tmp:
        csrc = os.path.join(tmp, "_prism.c")
        subprocess.check_call([sys.executable, "-m", "cython", "-3", REF_PYX,
                               "-o", csrc])
        subprocess.check_call([
            gcc, "-O2", "-fPIC", "-shared", "-w",
            "-I" + sysconfig.get_paths()["include"],
            "-I" + numpy.get_include(), csrc, "-o", os.path.join(tmp, "out.so"),
            "-lm"])
        shutil.move(os.path.join(tmp, "out.so"), out)
    return out


if __name__ == "__main__":
    path = build(force="--force" in sys.argv)
    print("oracle/_ref:", path if path else "reference tree absent, nothing built")
