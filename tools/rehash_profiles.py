"""
Re-key the committed kernel profiles without touching a measured number.

    python tools/rehash_profiles.py REV [key ...]

profiles/r02_kernel_<key>.json carry the hash of the sources their capture ran on.  The scheme
was refined twice in round 2: first ONE hash over all of csrc/ and include/ (`csrc_hash`), then
one hash per kernel over its translation unit, its includes AND the launch code, now two
hashes -- `kernel_hash` (device code of that kernel: bench.kernel_sources) and `launch_hash`
(prism_abi.cu, abi_internal.h, the public header: bench.launch_sources).  For each file (or
the named keys) this script checks out REV's csrc/ and include/ into a scratch directory,
VERIFIES that REV is the tree the capture ran on -- the recorded hash of whichever generation
the file carries must equal the same hash computed from REV's files -- and only then writes
the current pair of hashes computed from REV's files.  A file whose recorded hash does not
match REV is left alone.
"""
import glob
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def whole_tree_hash(root):
    """generation 1: everything under csrc/ plus include/*.h"""
    files = sorted(glob.glob(os.path.join(root, "fatiando_b200", "csrc", "*")) +
                   glob.glob(os.path.join(root, "include", "*.h")))
    return bench._hash_files(files)


def kernel_and_launch_hash(key, root):
    """generation 2: device sources and launch files in one sorted list, then the public header"""
    csrc = sorted(set(bench.kernel_sources(key, root)) |
                  set(os.path.join(root, "fatiando_b200", "csrc", f) for f in bench.LAUNCH_FILES))
    h = hashlib.sha256()
    for f in csrc + sorted(glob.glob(os.path.join(root, "include", "*.h"))):
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def main(rev, keys=()):
    short = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", rev], capture_output=True,
                           text=True).stdout.strip()
    with tempfile.TemporaryDirectory() as tmp:
        tar = subprocess.run(["git", "-C", ROOT, "archive", rev, "fatiando_b200/csrc", "include"],
                             check=True, capture_output=True).stdout
        subprocess.run(["tar", "-x", "-C", tmp], input=tar, check=True)
        for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_kernel_*.json"))):
            d = json.load(open(path))
            key = d["key"]
            if keys and key not in keys:
                continue
            name = os.path.basename(path)
            if "launch_hash" in d:
                ok, how = d["kernel_hash"] == bench.kernel_hash(key, tmp) and d["launch_hash"] == bench.launch_hash(tmp), "current"
            elif "csrc_hash" in d and d.get("rehashed_from") is None:
                ok, how = d["csrc_hash"] == whole_tree_hash(tmp), "whole-tree"
            elif "csrc_hash" in d:
                # generation 1 file already re-keyed to generation 2: both must agree with REV
                ok = d["csrc_hash"] == whole_tree_hash(tmp) and d["kernel_hash"] == kernel_and_launch_hash(key, tmp)
                how = "whole-tree + kernel-and-launch"
            else:
                ok, how = d["kernel_hash"] == kernel_and_launch_hash(key, tmp), "kernel-and-launch"
            if not ok:
                print("%s: its recorded %s hash is not %s's -- left alone" % (name, how, short))
                continue
            d["kernel_hash"] = bench.kernel_hash(key, tmp)
            d["kernel_sources"] = [os.path.basename(f) for f in bench.kernel_sources(key, tmp)]
            d["launch_hash"] = bench.launch_hash(tmp)
            d["launch_sources"] = [os.path.basename(f) for f in bench.launch_sources(tmp)]
            d["rehashed_from"] = "git %s (recorded %s hash verified against that revision)" % (short, how)
            json.dump(d, open(path, "w"), indent=1)
            prof, src = bench.load_profile(key)
            print("%s: kernel %s launch %s -- %s" % (name, d["kernel_hash"], d["launch_hash"],
                                                      ("counts valid; launch code " + prof["launch_code"]) if prof else src))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
