"""CPU suite: physical invariants of the oracle's restated stepSimulation (parity with pybullet itself is
UNPINNED -- the wheel is absent -- so the restatement is anchored on first principles here and on the
independent CUDA formulation in the GPU suite)."""
import ctypes as C

import numpy as np
import pytest

from oracle.oracle import OracleSim

MASS = 4.52   # rex.urdf: 1.2 + 2*0.05 + 4*(0.1 + 0.1 + 0.5 + 0.1 + 0.005)


def fresh(no_damping=False, dt=None, **kw):
    s = OracleSim(1, "walk", settle=False, target_position=2.0, backwards=False, **kw)
    if no_damping or dt:
        if no_damping:
            s.cfg.link_damping = 0.0
            s.model.root_mass = 0.0
            for a in range(3):
                s.model.root_inertia[a] = 0.0
        if dt:
            s.cfg.sim_dt = dt
        s.L.rexo_destroy(s.h)
        s.h = s.L.rexo_create(C.byref(s.model), C.byref(s.cfg))
    s.reset()
    return s


def randomize(s, rng, z=5.0):
    e = s.env(0)
    for j in range(12):
        e.q[j] += rng.normal() * 0.3
        e.qd[j] = rng.normal() * 3
    for a in range(3):
        e.linvel[a] = rng.normal()
        e.angvel[a] = rng.normal() * 3
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    for a in range(4):
        e.quat[a] = q[a]
    e.pos[2] = z
    return e


def test_total_mass_and_mass_matrix():
    s = fresh()
    e = randomize(s, np.random.default_rng(1))
    Minv = s.mass_matrix_inv()
    assert np.abs(Minv - Minv.T).max() < 1e-9
    M = np.linalg.inv(Minv)
    np.testing.assert_allclose(np.diag(M)[3:6], MASS, rtol=1e-9)
    assert np.linalg.eigvalsh(0.5 * (M + M.T)).min() > 0
    # kinetic energy from body twists equals 1/2 v^T M v for several velocity vectors
    rng = np.random.default_rng(2)
    for _ in range(5):
        for j in range(12):
            e.qd[j] = rng.normal() * 3
        for a in range(3):
            e.linvel[a] = rng.normal(); e.angvel[a] = rng.normal()
        v = np.concatenate([np.array(e.angvel), np.array(e.linvel), np.array(e.qd[:12])])
        ke = s.momentum()[0]
        assert abs(ke - 0.5 * v @ M @ v) < 1e-10 * max(1, ke)


def test_free_fall_and_aba_linearity():
    s = fresh()
    acc = s.aba(np.zeros(12))
    np.testing.assert_allclose(acc[:6], [0, 0, 0, 0, 0, -10.0], atol=1e-12)      # setGravity(0,0,-10)
    np.testing.assert_allclose(acc[6:], 0, atol=1e-10)                           # joints do not move in free fall
    e = randomize(s, np.random.default_rng(3))
    Minv = s.mass_matrix_inv()
    rng = np.random.default_rng(4)
    t0, t1 = rng.normal(size=12), rng.normal(size=12)
    np.testing.assert_allclose(s.aba(t1) - s.aba(t0), Minv[:, 6:] @ (t1 - t0), atol=1e-9)


def test_conservation_laws_in_flight():
    s = fresh(no_damping=True, dt=1e-4)
    randomize(s, np.random.default_rng(5))
    ke, lin, ang, com = s.momentum()
    E0, L0 = ke + MASS * 10 * com[2], ang - np.cross(com, lin)
    for _ in range(2000):
        s.physics_only(0, np.zeros(12))
    ke2, lin2, ang2, com2 = s.momentum()
    E1, L1 = ke2 + MASS * 10 * com2[2], ang2 - np.cross(com2, lin2)
    assert abs(E1 - E0) < 5e-3 * ke                       # symplectic Euler at dt=1e-4
    np.testing.assert_allclose(lin2[:2], lin[:2], atol=1e-3)
    assert abs((lin2[2] - lin[2]) + MASS * 10 * 0.2) < 1e-3
    np.testing.assert_allclose(L1, L0, atol=2e-4)


def test_static_stand_settles_on_four_toes():
    s = OracleSim(1, "walk", target_position=2.0, backwards=False)
    s.reset()
    e = s.env(0)
    st = s.state()
    assert 0.200 < st["pos"][2] < 0.210
    assert np.abs(st["qd"]).max() < 0.1 and np.abs(st["linvel"]).max() < 0.02
    toes = [2, 4, 6, 8]
    assert e.contact_mask == sum(1 << t for t in toes)      # only the four foot groups (toe hulls) touch
    assert e.limit_rows == 0 and 1 <= e.solver_iters <= 60
    q4 = st["quat"]; assert abs(q4[3]) > 0.9999


def test_friction_cone_holds_a_standing_robot_against_a_small_push():
    s = OracleSim(1, "walk", target_position=2.0, backwards=False)
    s.reset()
    e = s.env(0)
    x0 = e.pos[0]
    e.linvel[0] = 0.05
    stand = np.array([0., -0.88643435, 1.30197369] * 4)
    for _ in range(400):
        s.substep(0, stand)
    assert abs(e.pos[0] - x0) < 0.02 and abs(e.linvel[0]) < 0.02


def test_heightfield_flat_equals_plane():
    flat = np.zeros((1, 256, 256), np.float32)
    a = OracleSim(1, "walk", target_position=2.0, backwards=False)
    # terrain_full_toe: sample every toe hull vertex on the heightfield too (the product samples every 4th one there)
    b = OracleSim(1, "walk", terrain="random", fields=flat, target_position=2.0, backwards=False, terrain_full_toe=True)
    b.cfg.friction = a.cfg.friction
    b.L.rexo_destroy(b.h); b.h = b.L.rexo_create(C.byref(b.model), C.byref(b.cfg))
    a.reset(); b.reset()
    rng = np.random.default_rng(0)
    for _ in range(60):
        act = rng.uniform(-0.4, 0.4, (1, 2))
        oa, ra, da = a.step(act); ob, rb, db = b.step(act)
    np.testing.assert_allclose(a.state()["q"], b.state()["q"], atol=1e-12)
    np.testing.assert_allclose(a.state()["pos"], b.state()["pos"], atol=1e-12)


def test_random_terrain_runs_and_uses_the_field():
    s = OracleSim(2, "turn", terrain="random", nfields=2)
    s.reset()
    z = [s.env(i).pos[2] for i in range(2)]
    assert s.env(0).field_id != s.env(1).field_id
    for _ in range(40):
        s.step(np.zeros((2, 2)))
    assert all(np.isfinite(s.state(i)["q"]).all() for i in range(2))
    assert abs(z[0] - 0.21) < 1e-12                         # turn_env.py:157-159 re-places the base at z=0.21


def test_f32_build_tracks_f64_over_a_short_horizon():
    a = OracleSim(1, "walk", target_position=2.0, backwards=False)
    b = OracleSim(1, "walk", target_position=2.0, backwards=False, f32=True)
    a.reset(); b.reset()
    rng = np.random.default_rng(0)
    for _ in range(100):
        act = rng.uniform(-0.4, 0.4, (1, 2))
        a.step(act); b.step(act)
    assert np.abs(a.state()["q"] - b.state()["q"]).max() < 1e-3
    assert np.abs(a.state()["pos"] - b.state()["pos"]).max() < 1e-3


def test_wrappers_and_episode_limit():
    s = OracleSim(3, "walk", target_position=2.0, backwards=False, normalize=True, max_episode_steps=5)
    obs = s.reset()
    assert np.abs(obs).max() <= 1.0
    for k in range(5):
        obs, r, d = s.step(np.full((3, 2), 7.0))            # ClipAction clips to 1 before denormalising
        assert d.all() == (k == 4)
    assert np.abs(obs).max() <= 1.0
