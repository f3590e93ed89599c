"""sensor latency and noise (rex_gym/model/rex.py:717-769).

Latency is deterministic and pinned to goldens produced by the reference's own Rex.ReceiveObservation /
_GetDelayedObservation / _GetPDObservation (tools/gen_sensor_golden.py imports them from /root/reference and runs them
unmodified).  Noise replaces the unseeded np.random.normal by a counter-based generator, so it is checked for its
distribution and for the places it enters (observation, reward, termination), not sample by sample."""
import ctypes as C
import gzip
import json
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from script import scripted_state  # noqa: E402
from oracle.oracle import OracleSim, lib  # noqa: E402

G = json.load(gzip.open(os.path.join(os.path.dirname(__file__), "golden", "sensor_golden.json.gz"), "rt"))


def euler_to_quat(rpy):
    hr, hp, hy = rpy[0] / 2, rpy[1] / 2, rpy[2] / 2
    cr, sr, cp, sp, cy, sy = math.cos(hr), math.sin(hr), math.cos(hp), math.sin(hp), math.cos(hy), math.sin(hy)
    return [sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy]


def true_obs(k):
    pos, rpy, angvel, q, qd, tau = scripted_state(k, 0.3, 0.1, 0.2, 1, G["dt"])
    return np.array(list(q) + list(qd) + list(tau) + list(euler_to_quat(rpy)) + list(angvel))


def test_delayed_observation_matches_the_reference_history():
    """deque(maxlen=100).appendleft + the interpolated read, including latencies beyond the ring (oldest entry) and the
    `len == 1` special case of the first read."""
    s = OracleSim(1, "walk", "ik", settle=False, control_time_step=0.005, action_repeat=5)
    assert abs(s.cfg.sim_dt - G["dt"]) < 1e-15
    L = s.L
    out = np.zeros(43)
    for case in G["cases"]:
        L.rexo_sensor_clear(s.h, 0)
        for k in range(G["steps"]):
            o = np.ascontiguousarray(true_obs(k))
            L.rexo_sensor_push(s.h, 0, o.ctypes.data_as(C.c_void_p))
            L.rexo_sensor_delayed(s.h, 0, case["control_latency"], out.ctypes.data_as(C.c_void_p))
            np.testing.assert_allclose(out, case["control_observation"][k], rtol=0, atol=1e-12)
            L.rexo_sensor_delayed(s.h, 0, case["pd_latency"], out.ctypes.data_as(C.c_void_p))
            np.testing.assert_allclose(out[:24], case["pd_observation"][k], rtol=0, atol=1e-12)


def test_latency_zero_and_noise_zero_is_the_plain_path():
    kw = dict(target_position=2.0, backwards=False)
    a, b = OracleSim(4, "walk", "ik", **kw), OracleSim(4, "walk", "ik", control_latency=1e-12, **kw)   # sensor path on, zero delay
    oa, ob = a.reset(), b.reset()
    np.testing.assert_allclose(oa, ob, atol=1e-7)
    rng = np.random.default_rng(0)
    for _ in range(40):
        act = rng.uniform(-0.4, 0.4, (4, 2)).astype(np.float32)
        oa, ra, da = a.step(act); ob, rb, db = b.step(act)
        np.testing.assert_allclose(oa, ob, atol=2e-6); np.testing.assert_allclose(ra, rb, atol=2e-6); assert (da == db).all()


def test_control_latency_delays_the_observation_by_whole_sub_steps():
    """control_latency = 2 control steps: the observation returned by step k is the true observation of step k-2."""
    kw = dict(target_position=2.0, backwards=False)
    a, b = OracleSim(2, "walk", "ik", **kw), OracleSim(2, "walk", "ik", control_latency=0.010, **kw)
    a.reset(); b.reset()
    rng = np.random.default_rng(1)
    hist = []
    for k in range(30):
        act = rng.uniform(-0.4, 0.4, (2, 2)).astype(np.float32)
        oa, _, _ = a.step(act); ob, _, _ = b.step(act)
        hist.append(oa.copy())
        if k >= 2:
            np.testing.assert_allclose(ob, hist[k - 2], atol=1e-6)       # the dynamics do not depend on control_latency (walk-ik)


def test_pd_latency_feeds_the_motor_model_stale_joint_state():
    kw = dict(target_position=2.0, backwards=False)
    a, b = OracleSim(1, "walk", "ik", **kw), OracleSim(1, "walk", "ik", pd_latency=0.003, **kw)
    a.reset(); b.reset()
    act = np.zeros((1, 2), np.float32)
    for _ in range(20):
        a.step(act); b.step(act)
    assert np.abs(a.state(0)["q"] - b.state(0)["q"]).max() > 1e-4          # a 3 ms old PD observation changes the torques


def test_noise_generator_is_standard_normal_and_keyed():
    L = lib()
    x = np.array([L.rexo_noise(7, e, 1, st, 0, c) for e in range(40) for st in range(50) for c in range(5)])
    assert abs(x.mean()) < 0.03 and abs(x.std() - 1) < 0.03 and abs(((x - x.mean()) ** 3).mean()) < 0.1
    assert abs(np.corrcoef(x[:-1], x[1:])[0, 1]) < 0.03
    assert L.rexo_noise(7, 3, 1, 5, 0, 1) == L.rexo_noise(7, 3, 1, 5, 0, 1) != L.rexo_noise(7, 3, 1, 5, 1, 1)


def test_observation_noise_has_the_configured_spread():
    """SENSOR_NOISE_STDDEV order (rex.py:22): motor angle, motor velocity, motor torque, base rpy, base rpy rate."""
    sd = (0.02, 0.0, 0.0, 0.01, 0.5)
    kw = dict(target_position=2.0)
    a, b = OracleSim(64, "gallop", "ol", **kw), OracleSim(64, "gallop", "ol", observation_noise_stdev=sd, **kw)
    oa, ob = a.reset(), b.reset()
    d = ob.astype(np.float64) - oa
    assert abs(d[:, 0:2].std() - 0.01) < 0.003 and abs(d[:, 2:4].std() - 0.5) < 0.12 and abs(d[:, 4:].std() - 0.02) < 0.003
    act = np.zeros((64, 4), np.float32)
    oa, ra, _ = a.step(act); ob, rb, _ = b.step(act)
    d = ob.astype(np.float64) - oa
    assert abs(d[:, 0:2].std() - 0.01) < 0.003 and abs(d[:, 4:].std() - 0.02) < 0.003
    assert np.abs(ra - rb).max() < 0.01 and np.abs(ra - rb).max() > 0          # shake term sees the noisy orientation (rex_gym_env.py:530)
