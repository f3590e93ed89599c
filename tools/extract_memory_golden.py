#!/usr/bin/env python3
"""Recover REAL PyBullet trajectories from the reference's shipped PPO checkpoints and commit them as golden fixtures.

Every checkpoint under /root/reference/rex_gym/policies/<task>/<signal>/ stores, next to the network weights, the PPO
EpisodeMemory variables (rex_gym/agents/ppo/memory.py:32-45; algorithm.py:59-62): `memory/Variable_1..5` = the observations,
actions, action means, log-stddevs and rewards of the last `update_every` training episodes, recorded by the reference's
own RexGymEnv on pybullet==2.8.3 behind its training wrappers (ClipAction / RangeNormalize / ConvertTo32Bit,
rex_gym/agents/scripts/utility.py).  `EpisodeMemory.clear` only zeroes the length vector, so the rows survive in the file.

For the OPEN-LOOP tasks the motor command depends only on the action and the simulation clock (gallop_env.py:286-304,
walk_env.py:292-315, turn_env.py:271-311) -- no wall-clock gait phase -- so replaying the stored actions from the stored
initial observation is a like-for-like test of pybullet.stepSimulation + the motor model, the part of the hot path whose
parity could not be pinned any other way (pybullet is not installable here).

Output: tests/golden/pybullet_memory_golden.npz  (first STEPS[task] control steps of all 25 episodes per task; float32 as stored).
  <task>_<signal>_action [E][K][A]    policy output, before ClipAction / denormalisation (wrappers.py:218-265)
  <task>_<signal>_observ [E][K+1][O]  RangeNormalize'd observation BEFORE each step (row 0 = reset observation)
  <task>_<signal>_reward [E][K]
The reader is rex_gym_b200/agents/tf_checkpoint.py (pure Python; TensorFlow is not needed).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from rex_gym_b200.agents import tf_checkpoint as tfc  # noqa: E402

REF = "/root/reference/rex_gym/policies"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "pybullet_memory_golden.npz")
TASKS = [("gallop", "ol"), ("walk", "ol"), ("turn", "ol"), ("standup", "ol")]
EPISODES = 25                      # update_every = 25: every episode the memory holds
# control steps kept per episode: gallop 600 (no recorded episode reaches its goal before step 511; the target, hence the brake
# phase, is a per-episode random draw that was not recorded), walk 600 (the first goal is reached after step 400), standup: all
# 400, turn 160 (rows beyond `length` belong to older episodes)
STEPS = {"gallop": 600, "walk": 600, "turn": 160, "standup": 400}


def main():
    out = {}
    for task, sig in TASKS:
        prefix = tfc.latest_checkpoint(os.path.join(REF, task, sig))
        v = tfc.load_variables(prefix, ["memory/Variable_1", "memory/Variable_2", "memory/Variable_5"])
        ob, ac, rw = v["memory/Variable_1"], v["memory/Variable_2"], v["memory/Variable_5"]
        name = "%s_%s" % (task, sig)
        n = STEPS[task]
        out[name + "_action"] = ac[:EPISODES, :n].astype(np.float32)
        out[name + "_observ"] = ob[:EPISODES, :n + 1].astype(np.float32)
        out[name + "_reward"] = rw[:EPISODES, :n].astype(np.float32)
        print(name, os.path.basename(prefix), "memory", ob.shape, "->", out[name + "_observ"].shape)
    # walk-ol through goal, brake and standstill: 1200 steps of the four episodes with the nearest targets (1.0 ... 1.8 m; the
    # target itself is recovered from the recorded reward, tests/test_pybullet_goldens.py)
    v = tfc.load_variables(tfc.latest_checkpoint(os.path.join(REF, "walk", "ol")), ["memory/Variable_1", "memory/Variable_2", "memory/Variable_5"])
    rows = [0, 22, 1, 5]
    out["walk_ol_long_episodes"] = np.array(rows, np.int32)
    out["walk_ol_long_action"] = v["memory/Variable_2"][rows, :1200].astype(np.float32)
    out["walk_ol_long_observ"] = v["memory/Variable_1"][rows, :1201].astype(np.float32)
    out["walk_ol_long_reward"] = v["memory/Variable_5"][rows, :1200].astype(np.float32)
    # walk-ik: NOT replayable step by step (the gait phase ran on the wall clock, gait_planner.py:108-110), kept for the
    # statistical test of the wall-clock emulation (gait_clock_scale): 300 steps of 6 episodes
    v = tfc.load_variables(tfc.latest_checkpoint(os.path.join(REF, "walk", "ik")), ["memory/Variable_1", "memory/Variable_2", "memory/Variable_5"])
    out["walk_ik_action"] = v["memory/Variable_2"][:6, :300].astype(np.float32)
    out["walk_ik_observ"] = v["memory/Variable_1"][:6, :301].astype(np.float32)
    out["walk_ik_reward"] = v["memory/Variable_5"][:6, :300].astype(np.float32)      # carries the base x (forward term), see the tests
    # gallop-ik: same story (wall-clock gait), but the observation carries the 12 joint angles: 600 steps of 3 episodes
    v = tfc.load_variables(tfc.latest_checkpoint(os.path.join(REF, "gallop", "ik")), ["memory/Variable_1", "memory/Variable_2", "memory/Variable_5"])
    out["gallop_ik_action"] = v["memory/Variable_2"][:3, :600].astype(np.float32)
    out["gallop_ik_observ"] = v["memory/Variable_1"][:3, :601].astype(np.float32)
    out["gallop_ik_reward"] = v["memory/Variable_5"][:3, :600].astype(np.float32)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
