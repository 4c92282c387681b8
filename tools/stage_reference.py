#!/usr/bin/env python
"""Stage the UNMODIFIED reference sources of the hot path for the CPU baseline on the GPU box.

`gpurun` ships only /root/repo; /root/reference does not exist there.  BASELINE.md section 3 /
SURVEY section 7 step 0 reserve the git-ignored `baseline/_ref/` for the reference: this script copies
the four hot-path files (and the metrics callback that consumes their `info`, for the
conformance test) byte for byte into baseline/_ref/src/ and writes their sha256 next to them
(bench.py prints the hashes of what it timed).  Nothing under baseline/_ref is tracked by git or
imported by the product; oracle/ref_harness.py finds it ($SALP_REF_DIR, then baseline/_ref/src, then
/root/reference/src).

The sanctioned `pip install --target baseline/_ref /root/reference` cannot work for this reference:
its setup.py opens a README.md that does not exist, and `find_packages(where="src")` finds no package
(src/ has no __init__.py), see DESIGN.md section 5.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ("dynamics.py", "geometry.py", "robot.py", "salp_robot_env.py", "tensorboard_callback.py")


def stage(src="/root/reference/src", quiet=False):
    if not os.path.isfile(os.path.join(src, "salp_robot_env.py")):
        if not quiet:
            print(f"stage_reference: {src} not present; nothing staged")
        return None
    dst = os.path.join(ROOT, "baseline", "_ref", "src")
    os.makedirs(dst, exist_ok=True)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
        with open(os.path.join(dst, f), "rb") as fh:
            manifest[f] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "sha256": manifest}, fh, indent=1)
    if not quiet:
        print("staged", ", ".join(FILES), "->", dst)
    return dst


if __name__ == "__main__":
    stage(*(sys.argv[1:2]))
