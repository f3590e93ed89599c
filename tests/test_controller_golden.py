"""CPU: the oracle's controller half against golden vectors produced by the reference's own Python
modules (tools/gen_golden.py; fixtures in tests/golden/).  Tolerance 1e-12 abs: both sides are
float64 and follow the same operation order; libm differences only."""
import gzip
import json
import math
import os

import numpy as np
import pytest

from oracle import oracle as O
from script import scripted_state

GOLD = json.loads(gzip.open(os.path.join(os.path.dirname(__file__), "golden", "controller_golden.json.gz")).read())
TOL = 1e-12


def euler_to_quat(rpy):
    hr, hp, hy = rpy[0] / 2, rpy[1] / 2, rpy[2] / 2
    cr, sr, cp, sp, cy, sy = math.cos(hr), math.sin(hr), math.cos(hp), math.sin(hp), math.cos(hy), math.sin(hy)
    return [sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy]


def test_motor_model():
    for c in GOLD["motor"]:
        ta, to = O.motor_torque(c["cmd"], c["q"], c["qd"], c["qd"], c["kp"], c["kd"])
        np.testing.assert_allclose(ta, c["tau_act"], rtol=0, atol=TOL)
        np.testing.assert_allclose(to, c["tau_obs"], rtol=0, atol=TOL)


def test_motor_model_survey_vector():
    """SURVEY.md section 8(c) quoted vector."""
    stand = np.array([0., -0.88643435, 1.30197369] * 4)
    ta, to = O.motor_torque(stand, stand + np.linspace(-.5, .5, 12), np.linspace(-20, 20, 12), np.linspace(-20, 20, 12), 1.0, 0.02)
    np.testing.assert_allclose(ta, GOLD["motor_survey"]["tau_act"], atol=TOL)
    np.testing.assert_allclose(to, GOLD["motor_survey"]["tau_obs"], atol=TOL)
    np.testing.assert_allclose(ta[:6], [3.5, 3.5, 3.5, 3.5, 3.12566, 1.450792], atol=1e-6)


def test_ik():
    for c in GOLD["ik"]:
        out = O.ik_solve(c["rpy"], c["pos"], c["frames"])
        np.testing.assert_allclose(out, np.array(c["angles"]), rtol=0, atol=TOL)


def test_ik_survey_vector():
    fr = np.array([[0.115, -0.0925, -0.2], [0.115, 0.0925, -0.2], [-0.115, -0.0925, -0.2], [-0.115, 0.0925, -0.2]])
    out = O.ik_solve([0, 0, 0], [0.01, 0, 0], fr)
    np.testing.assert_allclose(out, np.tile([0, -0.870332, 1.307879], (4, 1)), atol=1e-6)


def test_gait_points():
    for c in GOLD["gait_points"]["swing"]:
        np.testing.assert_allclose(O.bezier_swing(c["phi"], c["v"], c["angle"], c["direction"]), c["xyz"], rtol=0, atol=TOL)
    for c in GOLD["gait_points"]["stance"]:
        np.testing.assert_allclose(O.stance(c["phi"], c["v"], c["angle"]), c["xyz"], rtol=0, atol=TOL)


@pytest.mark.parametrize("idx", range(3))
def test_gait_sequences(idx):
    seq = GOLD["gait"][idx]
    g = O.GaitState(seq["mode"])
    for s in seq["steps"]:
        fr = g.loop(s["t"], s["v"], s["angle"], s["w_rot"], s["T"], s["direction"])
        np.testing.assert_allclose(fr, np.array(s["frames"]), rtol=0, atol=TOL)
        assert abs(g.phi.value - s["phi"]) <= TOL and abs(g.alpha.value - s["alpha"]) <= TOL


@pytest.mark.parametrize("idx", range(len(GOLD["envs"])))
def test_env_signal_reward_termination_obs(idx):
    """<task>._transform_action_to_motor_command / _reward / _termination / _get_observation of the
    reference, replayed along the same scripted base trajectory."""
    g = GOLD["envs"][idx]
    sim = O.OracleSim(1, g["task"], g["signal"], settle=False,
                      target_position=g.get("target_position"), backwards=g.get("backwards", False),
                      target_orient=g.get("target_orient"), init_orient=g.get("init_orient"))
    sim.reset()
    e = sim.env(0)
    sc = g["script"]
    if g["task"] == "turn":
        assert bool(e.clockwise) == g["clockwise"]
    else:
        for a in range(4):
            e.quat[a] = [0, 0, 0, 1][a]
    e.pos[0], e.pos[1], e.pos[2] = 0.0, 0.0, 0.2
    for k, st in enumerate(g["steps"]):
        assert abs(e.step_counter * g["dt"] - st["t"]) < 1e-15
        cmd = sim.transform_action(st["action"])
        np.testing.assert_allclose(cmd, st["cmd"], rtol=0, atol=TOL, err_msg=f"step {k}")
        e.step_counter += g["repeat"]
        pos, rpy, angvel, q, qd, tau = scripted_state(k, sc["xspeed"], sc["yawspeed"], sc["yaw0"], g["repeat"], g["dt"], sc["tilt_at"])
        quat = euler_to_quat(rpy)
        for a in range(3):
            e.pos[a], e.angvel[a] = pos[a], angvel[a]
        for a in range(4):
            e.quat[a] = quat[a]
        for j in range(12):
            e.q[j], e.qd[j], e.tau_obs[j] = q[j], qd[j], tau[j]
        r, d, obs = sim.reward_done_obs()
        assert abs(r - st["reward"]) <= 1e-12 * max(1.0, abs(r)), f"reward step {k}"
        assert d == st["done"], f"done step {k}"
        np.testing.assert_allclose(obs, st["obs"], rtol=0, atol=1e-12, err_msg=f"obs step {k}")
        assert [e.goal_reached, e.stay_still, e.env_goal_reached] == st["flags"], f"flags step {k}"


# ---- RexPosesEnv (envs/gym/poses_env.py), fixture tests/golden/poses_golden.json.gz ------------------------
POSES = json.loads(gzip.open(os.path.join(os.path.dirname(__file__), "golden", "poses_golden.json.gz")).read())


def test_poses_reset_rotation():
    """deque rotation base_y, base_z, roll, pitch, yaw (rex_gym_env.py:259; poses_env.py:159-162): the constructor's own
    reset() consumes the first entry, so the k-th user-visible reset (reset_count k) lands on entry k mod 5."""
    names = ["base_y", "base_z", "roll", "pitch", "yaw"]
    sim = O.OracleSim(1, "poses", "ik", settle=False)
    lo_hi = {0: (-0.007, 0.007), 1: (-0.048, 0.021), 2: (-math.pi / 4, math.pi / 4), 3: (-math.pi / 4, math.pi / 4), 4: (-math.pi / 4, math.pi / 4)}
    for r in POSES["rotation"][1:]:
        sim.reset()
        e = sim.env(0)
        assert e.reset_count == r["reset_index"]
        assert names[e.next_pose] == r["next_pose"] and r["in_range"]
        lo, hi = lo_hi[e.next_pose]
        assert lo <= e.target_value <= hi


@pytest.mark.parametrize("idx", range(len(POSES["envs"])))
def test_poses_signal_reward_obs(idx):
    g = POSES["envs"][idx]
    a = g["args"]
    sim = O.OracleSim(1, "poses", "ik", settle=False, base_y=a["base_y"], base_z=a["base_z"], base_roll=a["base_roll"],
                      base_pitch=a["base_pitch"], base_yaw=a["base_yaw"])
    sim.reset()
    e = sim.env(0)
    assert e.next_pose == g["next_pose"] and e.target_value == g["target_value"]
    for k, st in enumerate(g["steps"]):
        assert abs(e.step_counter * g["dt"] - st["t"]) < 1e-15
        cmd = sim.transform_action(st["action"])
        np.testing.assert_allclose(cmd, st["cmd"], rtol=0, atol=TOL, err_msg=f"step {k}")
        e.step_counter += g["repeat"]
        pos, rpy, angvel, q, qd, tau = scripted_state(k, 0.0, 0.0, 0.0, g["repeat"], g["dt"], None)
        quat = euler_to_quat(rpy)
        for c in range(3):
            e.pos[c], e.angvel[c] = pos[c], angvel[c]
        for c in range(4):
            e.quat[c] = quat[c]
        r, d, obs = sim.reward_done_obs()
        assert r == st["reward"] == 1.0 and d == st["done"] is False
        np.testing.assert_allclose(obs, st["obs"], rtol=0, atol=1e-12, err_msg=f"obs step {k}")
