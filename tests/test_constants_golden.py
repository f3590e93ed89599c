"""The constant tables of the path against the reference's own module constants (tests/golden/constants_golden.json,
tools/gen_constants_golden.py): pose tables, motor order, motor-model tables -- where the oracle and the model tables can
be read from Python; the controller goldens pin the same numbers once more through the functions that use them."""
import json
import os

import numpy as np
import pytest

from oracle.oracle import OracleSim

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "constants_golden.json")))


@pytest.mark.parametrize("task,signal,mark,pose", [("walk", "ik", "base", "stand"), ("walk", "ol", "base", "stand_ol"),
                                                   ("gallop", "ik", "base", "stand"), ("gallop", "ol", "base", "stand_ol"),
                                                   ("turn", "ik", "base", "stand"), ("standup", "ol", "base", "rest_position"),
                                                   ("standup", "ol", "arm", "rest_position"), ("walk", "ik", "arm", "stand")])
def test_reset_pose_is_the_reference_pose_table(task, signal, mark, pose):
    """Rex.ResetPose (rex.py:343-408) puts the joints at INIT_POSES[pose] (+ the arm's rest pose); a reset without the hold
    leaves exactly that in the oracle."""
    kw = dict(target_position=2.0, backwards=False) if task == "walk" else {}
    s = OracleSim(1, task, signal, mark=mark, settle=2, **kw)
    s.reset()
    q = s.state(0)["q"]
    want = list(G["init_poses"][pose]) + (list(G["arm_poses"]["rest"]) if mark == "arm" else [])
    assert len(q) == G["motors_num"][mark]
    np.testing.assert_array_equal(q, want)


def test_motor_order_and_model_tables():
    from rex_gym_b200 import model_tables as MT
    root = os.path.join(os.path.dirname(__file__), "..", "rex_gym_b200", "model")
    for mark in ("base", "arm"):
        j = json.load(open(os.path.join(root, f"rex_{mark}.json")))
        assert j["motor_names"] == G["motor_names"][mark]               # FL, FR, RL, RR x (shoulder, leg, foot) [+ arm m1..m6]
    assert G["rex"]["OVERHEAT_SHUTDOWN_TORQUE"] == 2.45 and G["rex"]["OVERHEAT_SHUTDOWN_TIME"] == 1.0     # rexsim_kernel.cu motor block
    assert G["rex"]["SENSOR_NOISE_STDDEV"] == [0.0] * 5
    from rex_gym_b200.envs import batched_env as B
    assert list(B.SENSOR_NOISE_STDDEV) == G["rex"]["SENSOR_NOISE_STDDEV"] and B.OBSERVATION_EPS == G["env"]["OBSERVATION_EPS"]
    assert G["motor"]["current_table"] == [0, 10, 20, 30, 40, 50, 60] and G["motor"]["torque_table"] == [0, 1, 1.9, 2.45, 3.0, 3.25, 3.5]
    assert (G["motor"]["MOTOR_VOLTAGE"], G["motor"]["MOTOR_RESISTANCE"], G["motor"]["MOTOR_TORQUE_CONSTANT"]) == (32.0, 0.186, 0.0954)
    assert G["env"]["NUM_SIMULATION_ITERATION_STEPS"] == 300
    for task, rep in (("walk", 5), ("gallop", 6)):                       # numSolverIterations = 300 / action_repeat (rex_gym_env.py:184)
        s = OracleSim(1, task, "ik", settle=2, **(dict(target_position=2.0, backwards=False) if task == "walk" else {}))
        assert s.cfg.solver_iterations == 300 // rep and s.cfg.action_repeat == rep
    assert G["robot_init_position"]["plane"] == [0, 0, 0.21] == G["robot_init_position"]["random"]
    assert MT.LINK_DAMPING == 0.04
