"""
Give the committed kernel profiles their per-kernel source hash.

    python tools/rehash_profiles.py REV

The profiles/r02_kernel_<key>.json files were first written with ONE hash over all of csrc/
and include/ (`csrc_hash`), so a change to the polygonal-prism header made the profiles of the
unrelated prism kernels read as stale.  This script re-keys them without touching a measured
number: it checks out REV's csrc/ and include/ into a scratch directory, verifies that the
whole-tree hash of REV equals the `csrc_hash` each file recorded on the GPU box (so REV IS the
tree the capture ran on), and only then writes `kernel_hash` = bench.kernel_hash(key) computed
from REV's files, next to the original `csrc_hash`.  bench.load_profile compares `kernel_hash`
with the working tree's: a profile stays valid exactly while the sources of its own kernel do.
"""
import glob
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main(rev):
    with tempfile.TemporaryDirectory() as tmp:
        tar = subprocess.run(["git", "-C", ROOT, "archive", rev, "fatiando_b200/csrc", "include"],
                             check=True, capture_output=True).stdout
        subprocess.run(["tar", "-x", "-C", tmp], input=tar, check=True)
        saved, bench.ROOT = bench.ROOT, tmp
        whole = bench.csrc_hash()
        bench.ROOT = saved
        for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_kernel_*.json"))):
            d = json.load(open(path))
            key = d["key"]
            if d.get("csrc_hash") != whole:
                print("%s: recorded on csrc %s, %s is %s -- left alone" % (os.path.basename(path), d.get("csrc_hash"), rev, whole))
                continue
            d["kernel_hash"] = bench.kernel_hash(key, tmp)
            d["kernel_sources"] = [os.path.basename(f) for f in bench.kernel_sources(key, tmp)]
            d["rehashed_from"] = "git %s (whole-tree hash verified equal to csrc_hash)" % subprocess.run(
                ["git", "-C", ROOT, "rev-parse", "--short", rev], capture_output=True, text=True).stdout.strip()
            json.dump(d, open(path, "w"), indent=1)
            now = bench.kernel_hash(key)
            print("%s: kernel_hash %s (%s)" % (os.path.basename(path), d["kernel_hash"],
                                               "valid for the working tree" if now == d["kernel_hash"] else "STALE: working tree is " + now))


if __name__ == "__main__":
    main(sys.argv[1])
