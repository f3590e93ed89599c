"""Physics parity against REAL PyBullet trajectories.

tests/golden/pybullet_memory_golden.npz holds the first 160 control steps of 12 training episodes per open-loop task,
recovered from the PPO EpisodeMemory variables inside the checkpoints the reference ships (tools/extract_memory_golden.py):
policy actions, RangeNormalize'd observations and rewards recorded by the reference's own RexGymEnv on pybullet==2.8.3.
Open-loop motor commands depend only on the action and the simulation clock, so replaying the stored actions from the
stored reset observation tests pybullet.stepSimulation + the motor model like for like.

Every episode starts from the pristine pose (base at z = 0.21, joints exactly at the task's init pose, zero velocity: the
stored reset observation is exactly that), i.e. a 5 mm free fall onto the feet followed by hopping (gallop) or stepping
(walk).  Measured agreement of the fp64 oracle, median over the 12 episodes (tools/dev_pybullet_replay.py):
  gallop-ol  joint angles: 2.4e-4 rad after step 1, 6e-4 after 2 (free fall: ABA + motor model), 6e-3 through step 20
             (touchdown), 2.0e-2 rad worst sample within 150 steps (900 sub-steps of hopping), 5e-3 on average;
             pitch: 9e-3 rad worst sample, 3e-3 on average
  walk-ol    roll / pitch: 2.4e-3 rad worst sample within 150 steps
Legged contact is chaotic, so the bounds grow with the horizon; the thresholds below are ~1.7x the measured values.
The same file pins two modelling decisions that PyBullet's sources alone left open:
combined lateral friction 0.5 (0.25 / 1.0 are 5-7x worse) and no collision margin on the exact toe hull.
"""
import math
import os

import numpy as np
import pytest

from oracle.oracle import OracleSim

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "pybullet_memory_golden.npz"))
UA, UR = 2 * math.pi + 0.01, 2 * math.pi / 0.001 + 0.01       # RangeNormalize bounds (walk_env.py:364-374, rex_gym_env.py:277-278)
EPISODES, STEPS = 12, 150


def denorm(o):
    o = np.array(o, np.float64)
    o[..., 0:2] *= UA; o[..., 2:4] *= UR; o[..., 4:] *= UA
    return o


def replay_oracle(task, ep, steps=STEPS, **kw):
    name = task + "_ol"
    ac, ref = G[name + "_action"][ep], denorm(G[name + "_observ"][ep])
    s = OracleSim(1, task, "ol", normalize=True, settle=2, **kw)
    first = denorm(s.reset()[0])
    np.testing.assert_allclose(first, ref[0], atol=2e-5)          # the stored reset observation IS the pristine pose
    out = np.zeros((steps, ref.shape[1]))
    for t in range(steps):
        o, r, d = s.step(ac[t][None, :])
        out[t] = denorm(o[0])
        assert not d[0]
    return out, ref[1:steps + 1]


def errors(task, **kw):
    E = []
    for ep in range(EPISODES):
        ours, ref = replay_oracle(task, ep, **kw)
        rp = np.abs(ours[:, 0:2] - ref[:, 0:2]).max(1)
        q = np.abs(ours[:, 4:] - ref[:, 4:]).max(1) if ours.shape[1] > 4 else np.zeros(len(ours))
        E.append(np.stack([rp, q], 1))
    return np.array(E)                                            # [episode][step][rp, q]


def test_gallop_open_loop_tracks_pybullet():
    E = errors("gallop", target_position=2.0)
    q, rp = E[:, :, 1], E[:, :, 0]
    # free fall (no contact yet): articulated-body dynamics + motor model alone
    assert np.median(q[:, 0]) < 5e-4 and np.median(q[:, 1]) < 1.2e-3 and np.median(rp[:, 1]) < 1e-4
    # touchdown and the first hops
    assert np.median(q[:, :20].max(1)) < 1.1e-2 and np.median(rp[:, :20].max(1)) < 5e-3
    # 150 control steps = 900 physics sub-steps of hopping
    assert np.median(q.max(1)) < 3.5e-2 and np.median(rp.max(1)) < 1.6e-2
    assert np.median(q.mean(1)) < 8.5e-3 and np.median(rp.mean(1)) < 5e-3
    assert q.max() < 0.15                                         # no episode runs away


def test_walk_open_loop_tracks_pybullet():
    E = errors("walk", target_position=2.0, backwards=False)
    rp = E[:, :, 0]
    assert np.median(rp[:, :5].max(1)) < 5e-4
    assert np.median(rp.max(1)) < 4.5e-3 and rp.max() < 1.5e-2
    assert np.median(rp.mean(1)) < 1.5e-3


def _standup_replay(ep, steps):
    ac, ref, rw = G["standup_ol_action"][ep], denorm(G["standup_ol_observ"][ep]), G["standup_ol_reward"][ep]
    s = OracleSim(1, "standup", "ol", normalize=True)              # the full reset hold (rex.py:314-323), not the pristine pose
    s.reset()
    P, R = [], []
    for t in range(steps):
        o, r, _ = s.step(ac[t][None, :])
        P.append(denorm(o[0])[1]); R.append(r[0])
    return np.array(P), np.array(R), ref[1:steps + 1, 1], rw[:steps]


def test_standup_hop_tracks_pybullet_through_the_first_30_steps():
    """Standup episodes start from the belly-down rest pose the reset hold ends in, so they cannot be replayed from a pristine
    state; from our own hold the hop off the folded legs (first 30 control steps = 150 sub-steps, open loop) tracks the
    recorded pitch within 0.045 rad and crosses |pos - target| = 0.1 (reward sign flip) within one control step of PyBullet."""
    err, flip = [], []
    for ep in range(EPISODES):
        P, R, pref, rref = _standup_replay(ep, 30)
        err.append(np.abs(P - pref).max())
        flip.append(abs(int(np.argmax(R > 0)) - int(np.argmax(rref > 0))))
    assert np.median(err) < 0.045 and max(err) < 0.06
    assert max(flip) <= 1


@pytest.mark.xfail(reason="DESIGN.md section 9: PyBullet's reset hold ends belly-down (z ~ 0.039, feet past the 2.59 rad limit), "
                          "ours on the toes (z 0.0657); first reward -0.160 vs -0.184", strict=False)
def test_standup_rest_pose_matches_the_recorded_first_reward():
    _, R, _, rref = _standup_replay(0, 1)
    assert abs(R[0] - rref[0]) < 5e-3


def _with_cfg(task, ep, **ov):
    """Replay with overridden physics constants (friction etc.)."""
    import ctypes as C
    name = task + "_ol"
    ac, ref = G[name + "_action"][ep], denorm(G[name + "_observ"][ep])
    s = OracleSim(1, task, "ol", normalize=True, settle=2, target_position=2.0)
    for k, v in ov.items():
        setattr(s.cfg, k, v)
    s.L.rexo_destroy(s.h)
    s.h = s.L.rexo_create(C.byref(s.model), C.byref(s.cfg))
    s.reset()
    err = []
    for t in range(100):
        o, r, d = s.step(ac[t][None, :])
        err.append(np.abs(denorm(o[0])[4:] - ref[t + 1][4:]).max())
    return float(np.mean(err))


def test_recorded_trajectories_identify_the_friction_coefficient():
    """URDF link default 0.5 x plane.urdf lateral_friction 1.0 = 0.5; `<contact_coefficients mu="100">` (rex.urdf:194) is
    ignored by Bullet's URDF parser.  The recorded hopping is sharply selective: any other value is several times worse."""
    eps = range(6)
    base = np.median([_with_cfg("gallop", e) for e in eps])
    for mu in (0.25, 1.0):
        other = np.median([_with_cfg("gallop", e, friction=mu) for e in eps])
        assert other > 2.5 * base, (mu, other, base)


def test_fixture_matches_the_reference_checkpoints():
    """Regenerate-and-compare when the reference tree is present (it is not on the GPU box)."""
    ref = "/root/reference/rex_gym/policies"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present")
    from rex_gym_b200.agents import tf_checkpoint as tfc
    for task in ("gallop", "walk", "turn", "standup"):
        v = tfc.load_variables(tfc.latest_checkpoint(os.path.join(ref, task, "ol")), ["memory/Variable_1", "memory/Variable_2", "memory/Variable_5"])
        np.testing.assert_array_equal(G[task + "_ol_observ"], v["memory/Variable_1"][:12, :161])
        np.testing.assert_array_equal(G[task + "_ol_action"], v["memory/Variable_2"][:12, :160])
        np.testing.assert_array_equal(G[task + "_ol_reward"], v["memory/Variable_5"][:12, :160])


def test_wall_clock_gait_of_the_recorded_walk_ik_episodes():
    """GaitPlanner.loop reads time.time() (gait_planner.py:108-110).  The walk-ik episodes stored in the shipped checkpoint
    show what that meant in training: a 7.8-8.0 control-step ripple in roll/pitch, i.e. a gait cycle of ~8 control steps =
    40 ms of simulation time instead of the nominal 0.65 s -- the wall clock ran ~16x faster than the simulation.  With
    gait_clock_scale=16 the restatement reproduces the recorded regime (a level, shuffling walk: roll/pitch ripple of a few
    mrad for the whole window); with the simulation clock (scale 1) the same actions give the nominal trot, which is a
    different motion altogether (and tips over in this model)."""
    ref = denorm(G["walk_ik_observ"])
    ac = G["walk_ik_action"]
    np.testing.assert_array_equal(ref[:, 0], 0.0)                       # pristine reset observation
    for ep in range(3):
        seg = ref[ep, 100:300, 1] - ref[ep, 100:300, 1].mean()
        spec = np.abs(np.fft.rfft(seg)); k = int(np.argmax(spec[10:])) + 10   # ignore the slow drift (periods > 20 steps)
        assert 7.0 < len(seg) / k < 9.0                                 # recorded ripple period in control steps
        s = OracleSim(1, "walk", "ik", normalize=True, settle=2, target_position=2.0, backwards=False, gait_clock_scale=16.0)
        s.reset()
        out = []
        for t in range(300):
            o, r, d = s.step(ac[ep, t][None, :])
            assert not d[0], (ep, t)
            out.append(denorm(o[0]))
        out = np.array(out)
        for j in (0, 1):                                                # roll, pitch ripple: same size as recorded
            ours, real = out[50:, j].std(), ref[ep, 51:301, j].std()
            assert 0.4 * real < ours < 2.5 * real, (ep, j, ours, real)
        assert np.abs(out[:, :2]).max() < 0.04                          # level for the whole window (recorded: < 0.02)
    s = OracleSim(1, "walk", "ik", normalize=True, settle=2, target_position=2.0, backwards=False)      # simulation clock
    s.reset()
    big = 0.0
    for t in range(200):
        o, r, d = s.step(ac[0, t][None, :])
        big = max(big, float(np.abs(denorm(o[0])[:2]).max()))
        if d[0]:
            break
    assert big > 0.1                                                     # the nominal trot rocks the base 25x more


def test_wall_clock_gait_of_the_recorded_gallop_ik_episodes():
    """Same wall-clock effect on the gallop-ik policy's training data, where the observation also carries the 12 joint angles:
    the recorded leg-joint ripple has a 7.0 control-step period (0.3 s gait period / 42 ms => the wall clock ran ~7.1x
    faster than the simulation).  At gait_clock_scale = 8.5 (7-step cycle) the restatement reproduces period and size of the
    joint-angle ripple (leg 0.059 rad, foot 0.046 rad recorded) -- IK-driven legs through the motor model at 24 Hz."""
    ref = denorm(G["gallop_ik_observ"])
    ac = G["gallop_ik_action"]
    scale = 8.5      # phase step 6 ms * 8.5 / 0.3 s = 0.17 per control step: phi >= 0.99 after 6 steps, +1 step for the restart = 7
    for ep in range(3):
        s = OracleSim(1, "gallop", "ik", normalize=True, settle=2, target_position=2.0, gait_clock_scale=scale)
        s.reset()
        out = []
        for t in range(400):
            o, r, d = s.step(ac[ep, t][None, :])
            assert not d[0], (ep, t)
            out.append(denorm(o[0]))
        out = np.array(out)

        def period(x):
            x = x - x.mean()
            spec = np.abs(np.fft.rfft(x)); k = int(np.argmax(spec[5:])) + 5
            return len(x) / k
        assert abs(period(ref[ep, 101:401, 5]) - 7.0) < 0.5 and abs(period(out[100:, 5]) - 7.0) < 1.0
        for j in (5, 6, 11, 12):                                       # leg / foot joints of a front and a rear leg
            ours, real = out[100:, j].std(), ref[ep, 101:401, j].std()
            assert 0.6 * real < ours < 1.6 * real, (ep, j, ours, real)
        assert abs(out[100:, 1].mean() - ref[ep, 101:401, 1].mean()) < 0.03      # mean pitch of the hopping gait


@pytest.mark.skipif(not os.path.isdir("/root/reference/rex_gym/policies"), reason="needs the shipped TF checkpoint (reference tree)")
def test_shipped_gallop_policy_closes_the_loop_like_its_training_runs():
    """End to end: the gallop-ol policy the reference ships (TF checkpoint -> pure-Python reader -> numpy restatement of
    ForwardGaussianPolicy + StreamingNormalize) drives the restated simulator.  It gallops forward (-x) for the whole
    1000-step episode, and its progress is consistent with the rewards PyBullet produced during training: the forward
    reward is x / target with target ~ U(1, 3) (gallop_env.py:151, rex_gym_env.py:513-520), so every recorded reward implies
    a target x_ours(t) / reward(t), which has to fall in that range if the two simulators cover ground at the same rate."""
    from oracle import agent_oracle as AO
    from rex_gym_b200.agents.networks import read_tf_policy
    w, filt = read_tf_policy("/root/reference/rex_gym/policies/gallop/ol")
    f = AO.StreamingNormalize((16,), True, True, 5)
    f.count, f.mean, f.var_sum = filt[0], np.array(filt[1], np.float64), np.array(filt[2], np.float64)
    s = OracleSim(1, "gallop", "ol", normalize=True, settle=2, target_position=2.0)
    obs = s.reset()
    e = s.env(0)
    xs = []
    for t in range(1000):
        _, mean, _, _ = AO.perform(w, f, obs.astype(np.float64), False)
        obs, r, d = s.step(mean.astype(np.float32))
        xs.append(e.pos[0])
        assert not d[0], t
    xs = -np.array(xs)
    assert 0.75 < xs[499] < 1.3 and 1.6 < xs[999] < 2.6 and abs(e.pos[1]) < 0.3          # ~0.35 m/s, straight
    implied = xs[149] / G["gallop_ol_reward"][:, 149]                                       # recorded: sampled actions, 12 episodes
    assert np.sum((implied > 0.85) & (implied < 3.5)) >= 10, implied


@pytest.mark.skipif(not os.path.isdir("/root/reference/rex_gym/policies"), reason="needs the shipped TF checkpoint (reference tree)")
def test_shipped_walk_policy_reproduces_the_recorded_reward_profile():
    """The shipped walk-ol policy in closed loop on the restated simulator, target 2.0 m: ramps up, walks to the goal, brakes
    and stands still -- the per-step reward (rex_gym_env.py:501-542) has the profile of the 25 episodes PyBullet recorded in
    the checkpoint: r(400)/r(200) in [3.8, 4.6] (acceleration phase), a final plateau of -0.28 ... -0.41 (the robot needs
    0.3-0.4 m to stop past the goal: `tp - cx` once cx > tp + 0.15), and standing still for the rest of the 2000 steps."""
    from oracle import agent_oracle as AO
    from rex_gym_b200.agents.networks import read_tf_policy
    w, filt = read_tf_policy("/root/reference/rex_gym/policies/walk/ol")
    f = AO.StreamingNormalize((4,), True, True, 5)
    f.count, f.mean, f.var_sum = filt[0], np.array(filt[1], np.float64), np.array(filt[2], np.float64)
    s = OracleSim(1, "walk", "ol", normalize=True, settle=2, target_position=2.0, backwards=False)
    obs = s.reset()
    e = s.env(0)
    R, X = [], []
    for t in range(2000):
        _, mean, _, _ = AO.perform(w, f, obs.astype(np.float64), False)
        obs, r, d = s.step(mean.astype(np.float32))
        R.append(float(r[0])); X.append(e.pos[0])
        assert not d[0], t
    R, X = np.array(R), -np.array(X)
    assert 3.6 < R[400] / R[200] < 4.8                       # recorded: 3.76 ... 4.57
    assert 0.9 < X[400] < 1.2                                # recorded r(400) = x / target, e.g. 0.96 at target ~1.1
    assert -0.45 < R[-1] < -0.25 and abs(R[-1] - R[1200]) < 0.03      # recorded plateau: -0.414 ... -0.278, reached by step ~800
    assert abs(X[-1] - X[1200]) < 0.02 and 2.2 < X[-1] < 2.5          # stands still past the goal
