#!/usr/bin/env python3
"""Golden vectors for SURVEY row a9: the reference's own `Rex.Step` loop (rex_gym/model/rex.py:158-163) --
`ApplyAction` (:568-641: PD observation, MotorModel.convert_to_torque, overheat protection, TORQUE_CONTROL writes),
`stepSimulation`, `ReceiveObservation` (:726-733) -- imported from /root/reference and run UNMODIFIED on a Rex object created
with object.__new__.  The pybullet client is a recorder: `getJointState` / base getters return the scripted state of
tests/golden/script.py at the current sub-step, `stepSimulation` advances that sub-step, `setJointMotorControl2` stores the
torque it is handed.  So what is pinned is everything around the physics step: which torque reaches which joint, the overheat
counters (more than 1000 consecutive sub-steps above 2.45 N m, rex.py:13-14,601-608) and the motors they switch off.

Output: tests/golden/apply_action_golden.json.gz
"""
import collections
import gzip
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "tests", "golden"))
from gen_golden import install_stubs  # noqa: E402
from script import OVERHEAT_SUBSTEPS, OVERHEAT_REPEAT, overheat_joint_state, overheat_command  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "apply_action_golden.json.gz")
DT, NM = 0.001, 12


def stored(k):
    """Sub-steps whose torques are kept in the fixture (all counters / enabled flags are kept)."""
    return k < 30 or 990 <= k < 1020 or 1995 <= k < 2015 or k % 25 == 0


def main():
    install_stubs()
    from rex_gym.model import rex as rexmod, motor
    k = [0]
    torques = {}

    class Client:
        TORQUE_CONTROL = 99

        def getJointState(self, body, joint):
            q, qd = overheat_joint_state(k[0])
            return (q[joint], qd[joint], (0,) * 6, 0.0)

        def getBasePositionAndOrientation(self, body):
            return (0.0, 0.0, 0.2), (0.0, 0.0, 0.0, 1.0)

        def getBaseVelocity(self, body):
            return (0.0, 0.0, 0.0), (0.0, 0.0, 0.0)

        def setJointMotorControl2(self, bodyIndex, jointIndex, controlMode, force):
            assert controlMode == self.TORQUE_CONTROL
            torques[jointIndex] = float(force)

        def stepSimulation(self):
            k[0] += 1

    r = object.__new__(rexmod.Rex)
    r._pybullet_client, r.quadruped = Client(), 1
    r.num_motors, r.time_step, r._action_repeat = NM, DT, OVERHEAT_REPEAT
    r._motor_velocity_limit = np.inf
    r._kp, r._kd = 1.0, 0.02                                  # walk_env.py:39-40
    r._accurate_motor_model_enabled, r._pd_control_enabled, r._motor_overheat_protection = True, False, True
    r._torque_control_enabled = False
    r._motor_model = motor.MotorModel(motors_num=NM, torque_control_enabled=False, kp=r._kp, kd=r._kd)     # rex.py:136-139
    r._motor_direction = [1 for _ in range(NM)]               # rex.py:113
    r._motor_id_list = list(range(NM))
    r._observed_motor_torques = np.zeros(NM)
    r._overheat_counter = np.zeros(NM)                        # Reset, rex.py:301-303
    r._motor_enabled_list = [True] * NM
    r._step_counter = 0
    r._observation_history = collections.deque(maxlen=100)
    r._control_latency = r._pd_latency = 0.0
    r._observation_noise_stdev = (0.0,) * 5
    r.ReceiveObservation()                                    # Reset ends with one (rex.py:324 region)
    out = {"source": "rex_gym/model/rex.py:158-163,568-641,726-733 run unmodified; pybullet client = scripted recorder",
           "dt": DT, "action_repeat": OVERHEAT_REPEAT, "substeps": OVERHEAT_SUBSTEPS, "kp": r._kp, "kd": r._kd,
           "stored_substeps": [], "applied": [], "observed": [], "counter": [], "enabled": []}
    orig = rexmod.Rex.ApplyAction

    def recording_apply(self, cmd, *a, **kw):                  # wraps, does not replace: the reference body runs as is
        orig(self, cmd, *a, **kw)
        kk = k[0]
        if stored(kk):
            out["stored_substeps"].append(kk)
            out["applied"].append([torques[j] for j in range(NM)])
            out["observed"].append([float(x) for x in self._observed_motor_torques])
        out["counter"].append([int(x) for x in self._overheat_counter])
        out["enabled"].append([int(bool(x)) for x in self._motor_enabled_list])
    rexmod.Rex.ApplyAction = recording_apply
    for c in range(OVERHEAT_SUBSTEPS // OVERHEAT_REPEAT):
        r.Step(overheat_command(c))
    assert k[0] == OVERHEAT_SUBSTEPS and r._step_counter == OVERHEAT_SUBSTEPS
    en = np.array(out["enabled"])
    first_off = [int(np.argmin(en[:, j])) if not en[:, j].all() else -1 for j in range(NM)]
    out["first_disabled_substep"] = first_off
    print("first sub-step with the motor off:", first_off)
    print("motor 3 counter max", max(c[3] for c in out["counter"]))
    with gzip.open(OUT, "wt") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(out["stored_substeps"]), "torque rows")


if __name__ == "__main__":
    main()
