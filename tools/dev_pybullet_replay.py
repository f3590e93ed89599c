#!/usr/bin/env python3
"""Developer tool: replay the PyBullet trajectories of tests/golden/pybullet_memory_golden.npz through the fp64 oracle and
print the error growth per control step (the numbers behind tests/test_pybullet_goldens.py)."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.oracle import OracleSim  # noqa: E402

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "pybullet_memory_golden.npz"))
UA, UR = 2 * math.pi + 0.01, 2 * math.pi / 0.001 + 0.01


def denorm(o):
    o = np.array(o, np.float64)
    o[..., 0:2] *= UA; o[..., 2:4] *= UR; o[..., 4:] *= UA
    return o


def replay(task, sig, ep, steps, **kw):
    name = "%s_%s" % (task, sig)
    ac, ob = G[name + "_action"][ep], denorm(G[name + "_observ"][ep])
    s = OracleSim(1, task, sig, normalize=True, settle=2, **kw)
    s.reset()
    out = []
    for t in range(steps):
        o, r, d = s.step(ac[t][None, :])
        out.append(denorm(o[0]))
    return np.array(out), ob[1:steps + 1], s


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    neps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    marks = [1, 2, 3, 5, 10, 20, 40, 80, 120, 150]
    for task, sig, kw in (("gallop", "ol", dict(target_position=2.0)), ("walk", "ol", dict(target_position=2.0, backwards=False))):
        E = []
        for ep in range(neps):
            ours, ref, _ = replay(task, sig, ep, steps, **kw)
            rp = np.abs(ours[:, 0:2] - ref[:, 0:2]).max(1)
            rate = np.abs(ours[:, 2:4] - ref[:, 2:4]).max(1)
            q = np.abs(ours[:, 4:] - ref[:, 4:]).max(1) if ours.shape[1] > 4 else np.zeros(steps)
            E.append(np.stack([rp, rate, q], 1))
        E = np.array(E)
        print(task, sig, "median / max over %d episodes of |roll,pitch| err, |rates| err, |joint| err at step:" % neps)
        for m in marks:
            if m <= steps:
                print("  step %4d  rp %.2e / %.2e   rate %.2e / %.2e   q %.2e / %.2e" % (
                    m, np.median(E[:, m - 1, 0]), E[:, m - 1, 0].max(), np.median(E[:, m - 1, 1]), E[:, m - 1, 1].max(),
                    np.median(E[:, m - 1, 2]), E[:, m - 1, 2].max()))
        print("  cumulative max up to step: " + "  ".join("%d: rp %.1e q %.1e" % (m, np.median(E[:, :m, 0].max(1)), np.median(E[:, :m, 2].max(1))) for m in marks if m <= steps))


if __name__ == "__main__":
    main()
