#!/usr/bin/env python3
"""Sub-step level GPU-vs-oracle trace (action_repeat=1): where do isolated parity events come from?"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from oracle.oracle import OracleSim
import rex_gym_b200 as R

n = 8
kw = dict(target_position=3.0, backwards=True, action_repeat=1, control_time_step=0.001)
env = R.BatchedRexEnv(task="walk", num_envs=n, solver_iterations=60, **kw)
ora = OracleSim(n, "walk", "ik", solver_iterations=60, **kw)
env.reset(); ora.reset()
rng = np.random.default_rng(5)
prev = None
events = 0
hist = []
for k in range(6000):
    if k % 125 == 0:
        so = [ora.state(i) for i in range(n)]
        S = {key: np.stack([s[key] for s in so]) for key in so[0]}
        env.set_state(S["pos"], S["quat"], S["linvel"], S["angvel"], S["q"], S["qd"])
        prev = np.zeros(n)
    a = rng.uniform(-0.4, 0.4, size=(n, 2)).astype(np.float32)
    env.step(a); ora.step(a)
    sg = env.get_state()
    dq = np.stack([sg["q"][i] - ora.state(i)["q"] for i in range(n)])
    dqd = np.stack([sg["qd"][i] - ora.state(i)["qd"] for i in range(n)])
    eq = np.abs(dq).max(axis=1)
    cm = np.array([ora.env(i).contact_mask & 0x1FF for i in range(n)])
    it = np.array([ora.env(i).solver_iters for i in range(n)])
    hist.append((eq.copy(), np.abs(dqd).max(axis=1), cm, sg["contact_mask"].copy(), it))
    jump = np.nonzero((eq > 1e-4) & (prev < 2e-5))[0]
    for i in jump:
        events += 1
        if events <= 4:
            print(f"event env {i} sub-step {k}: |dq| {prev[i]:.1e} -> {eq[i]:.1e}, joint {np.abs(dq[i]).argmax()}")
            for kk in range(max(0, k - 6), k + 1):
                h = hist[kk]
                print(f"   {kk}: dq {h[0][i]:.2e} dqd {h[1][i]:.2e} masks o/g {h[2][i]:09b}/{h[3][i]:09b} iters {h[4][i]}")
    prev = eq
print("events:", events)
e = np.stack([h[0] for h in hist])
print("percentiles of |dq| (50,90,95,99,max):", [float(np.percentile(e, p)) for p in (50, 90, 95, 99, 100)])
