#!/usr/bin/env python
"""Evidence for DESIGN.md 3.4: the reference does not reproduce ITSELF in the tumbling regime.

Runs the live, unmodified Python reference twice on the actions of one 500-cycle episode of
tests/golden/ref_long.npz -- once as recorded, once with the initial roll angle moved by 1e-15 rad
(the size of ONE float64 rounding of an O(1) quantity; world position and yaw would not do, the
body-frame dynamics are invariant under them) -- and prints how far the two runs are apart as the episode proceeds,
next to the golden max(|roll|, |pitch|).  Build container only (needs the reference).
    python tools/ref_self_divergence.py [env index]   -> profiles/r02_reference_self_divergence.json
"""
import json
import os
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_harness as rh  # noqa: E402


def run(args):
    actions, target, obstacles, eps = args
    env = rh.make_env()
    env.reset()
    rh.inject_scene(env, target, obstacles)
    env.robot.euler_angle[0] += eps
    out = []
    for a in actions:
        env.step(a.copy())
        r = env.robot
        out.append([r.position_world[0], r.position_world[1], r.euler_angle[0], r.euler_angle[1], r.euler_angle[2]])
    return np.array(out)


def main():
    i = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_long.npz"))
    acts = g["actions"][i][:500]
    t, o = g["targets"][i, 0], g["obstacles"][i, 0]
    with Pool(2) as pool:
        a, b = pool.map(run, [(acts, t, o, 0.0), (acts, t, o, 1e-15)])
    d_pos = np.hypot(a[:, 0] - b[:, 0], a[:, 1] - b[:, 1])
    d_yaw = np.abs(a[:, 4] - b[:, 4])
    tilt = np.maximum(np.abs(a[:, 2]), np.abs(a[:, 3]))
    rows = [dict(cycle=int(k + 1), tilt_rad=float(tilt[k]), position_gap_m=float(d_pos[k]), yaw_gap_rad=float(d_yaw[k]))
            for k in list(range(9, 500, 10))]
    out = dict(note="live Python reference vs itself, initial roll moved by 1e-15 rad; ref_long.npz env %d" % i, rows=rows)
    path = os.path.join(ROOT, "profiles", "r02_reference_self_divergence.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    for r in rows[::5]:
        print(r)
    for thr in (1e-9, 1e-3, 1.0):
        first = next((r["cycle"] for r in rows if r["position_gap_m"] > thr), None)
        out[f"first_cycle_with_position_gap_above_{thr:g}_m"] = first
        print(f"first recorded cycle with a position gap > {thr:g} m:", first)
    with open(path, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
