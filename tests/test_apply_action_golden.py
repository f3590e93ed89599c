"""SURVEY row a9 pinned to the reference's own `Rex.Step` / `Rex.ApplyAction` (rex_gym/model/rex.py:158-163,568-641).

tests/golden/apply_action_golden.json.gz (tools/gen_apply_action_golden.py) holds what the unmodified reference methods did over
2300 sub-steps of a scripted joint trajectory (tests/golden/script.py) with a recording pybullet client: the torque written to
every joint, the observed torque, the overheat counters and the enabled flags.  The oracle's `apply_action` (the part of
`apply_action_and_step` before the physics step; the CUDA path is compared with the oracle on counters and flags in
tests/test_gpu_parity.py) is driven through the same trajectory."""
import gzip
import json
import os
import sys

import numpy as np

from oracle.oracle import OracleSim

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from script import OVERHEAT_REPEAT, OVERHEAT_SUBSTEPS, overheat_command, overheat_joint_state  # noqa: E402

G = json.load(gzip.open(os.path.join(os.path.dirname(__file__), "golden", "apply_action_golden.json.gz"), "rt"))


def test_apply_action_torques_counters_and_shutdown_follow_the_reference():
    assert G["substeps"] == OVERHEAT_SUBSTEPS and G["action_repeat"] == OVERHEAT_REPEAT
    s = OracleSim(1, "walk", "ik", settle=2, motor_kp=G["kp"], motor_kd=G["kd"], target_position=2.0, backwards=False)
    s.reset()
    assert s.cfg.sim_dt == G["dt"] and s.cfg.action_repeat == OVERHEAT_REPEAT
    e = s.env(0)
    stored = {k: i for i, k in enumerate(G["stored_substeps"])}
    counter, enabled = np.array(G["counter"]), np.array(G["enabled"])
    worst_a = worst_o = 0.0
    for k in range(OVERHEAT_SUBSTEPS):
        q, qd = overheat_joint_state(k)
        for j in range(12):
            e.q[j] = q[j]; e.qd[j] = qd[j]                  # motor_dof is the identity on the base mark
        tau = s.apply_action(0, overheat_command(k // OVERHEAT_REPEAT))
        np.testing.assert_array_equal(np.array(e.overheat[:12], np.int64), counter[k], err_msg=str(k))
        np.testing.assert_array_equal(np.array(e.enabled[:12], np.int64), enabled[k], err_msg=str(k))
        if k in stored:
            worst_a = max(worst_a, np.abs(tau - G["applied"][stored[k]]).max())
            worst_o = max(worst_o, np.abs(np.array(e.tau_obs[:12]) - G["observed"][stored[k]]).max())
    assert worst_a < 1e-12 and worst_o < 1e-12, (worst_a, worst_o)
    # the script's design points, as the reference resolved them: off on the 1001st consecutive hot sub-step, one relieved
    # control step restarts the count, a shut-down motor stays off when the load goes away, the sign does not matter
    assert G["first_disabled_substep"][:5] == [1000, 2005, 1000, -1, 1000]
    k_off = G["stored_substeps"].index(1000)
    assert G["applied"][k_off][0] == 0.0 and abs(G["applied"][k_off - 1][0]) > 2.45 and abs(G["observed"][k_off][0]) > 2.45
    assert enabled[-1, 2] == 0 and counter[-1, 2] == 0      # relieved after shutdown: counter back to 0, motor still off
    assert 0 < counter[:, 3].max() < 1000 and (np.diff(counter[:, 3]) < 0).sum() >= 3     # motor 3 keeps crossing the threshold
