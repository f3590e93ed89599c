#!/usr/bin/env python3
"""Developer experiment (CPU, oracle only): what is the reference's rest pose at the start of a RexStandupEnv episode?

Replays the recorded PyBullet standup episodes of tests/golden/pybullet_memory_golden.npz through the fp64 oracle and prints the
rest state, the first reward, the reward increments of the first steps and the pitch trace next to the recorded ones.

With the oracle patched by tools/experiments/standup_rest_oracle.patch (NOT applied in the tree: `patch -p0 < ...` in a scratch
copy, `make -C oracle`), three environment knobs become available:

  REXO_VMAX=100            clamp every generalised velocity to +-100 (btMultiBody::m_maxCoordinateVelocity)
  REXO_PHI_F / REXO_PHI_R  freeze the front / rear foot joints at this angle from sub-step 12 of the 500-step hold on
  REXO_TRACE=1             per-sub-step trace of the hold phase on stderr
  (the patch also drops the positional correction of a joint-limit row once the violation exceeds 0.04 rad, the
   split-impulse branch of btMultiBodyJointLimitConstraint::createConstraintRows)

Findings (round 1):
  * recorded: first reward -0.1848 and reward 0.990 -> -1.989 when z crosses 0.21 at step 37/38, i.e. |x|+|y| ~ 0.010 there and
    z_rest ~ 0.038-0.040: the robot lies on its shoulder boxes / belly; pitch rate at reset -0.047 rad/s (still settling).
  * the tree's restatement rests on its toes at z = 0.0657 (first reward -0.160) because the foot joints stop at the 2.59 rad limit.
  * feet frozen anywhere in 2.9 .. 5.0 rad reproduce z = 0.0395 and the first reward (-0.184).
  * after the 100-step stand phase the foot motors whip the feet to the limit at 145-175 rad/s (100 with the clamp), so the
    joints overshoot the limit by up to 0.1 rad in one sub-step; which side of the 0.04 rad split-impulse threshold each leg
    lands on decides the rest pose.  Clamp + split give 2.69-2.72 rad and a still-creeping reset (pitch rate -0.058), not
    yet the recorded belly-down pose.
"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle.oracle import OracleSim  # noqa: E402

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests", "golden", "pybullet_memory_golden.npz"))
UA, UR = 2 * math.pi + 0.01, 2 * math.pi / 0.001 + 0.01


def main():
    ep = int(os.environ.get("EP", "0"))
    steps = int(os.environ.get("T", "60"))
    ac = G["standup_ol_action"][ep]
    ob = G["standup_ol_observ"][ep].astype(np.float64)
    rw = G["standup_ol_reward"][ep]
    ob[:, 0:2] *= UA
    ob[:, 2:4] *= UR
    kw = {}
    if os.environ.get("IT"):
        kw["solver_iterations"] = int(os.environ["IT"])
    s = OracleSim(1, "standup", "ol", normalize=True, **kw)     # normalize=True: the recorded actions are pre-RangeNormalize
    s.reset()
    st = s.state()
    R, P = [], []
    for t in range(steps):
        o, r, _ = s.step(ac[t][None, :])
        R.append(r[0])
        P.append(o[0, 1] * UA)
    R, P = np.array(R), np.array(P)
    err = np.abs(P - ob[1:steps + 1, 1])
    print("rest: z %.4f x %.4f pitch %.4f pitch-rate %.3f  feet %.3f / %.3f   (recorded pitch %.4f rate %.3f)" % (
        st["pos"][2], st["pos"][0], 2 * st["quat"][1], st["angvel"][1], st["q"][2], st["q"][8], ob[0, 1], ob[0, 3]))
    print("first reward %.4f (recorded %.4f); first positive reward at step %d (recorded %d)" % (
        R[0], rw[0], int(np.argmax(R > 0)) + 1, int(np.argmax(rw > 0)) + 1))
    print("pitch error over %d steps: mean %.4f max %.4f" % (steps, err.mean(), err.max()))
    print("reward increments [1e-3/step]  ours", np.round(np.diff(R[:10]) * 1e3, 1), " recorded", np.round(np.diff(rw[:10]) * 1e3, 1))
    print("pitch every 2nd step  ours", np.round(P[:24:2], 3))
    print("                  recorded", np.round(ob[1:25:2, 1], 3))


if __name__ == "__main__":
    main()
