#!/usr/bin/env python3
"""Developer probe (GPU box): CUDA path vs fp64 oracle on identical action sequences; prints error growth."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import torch
from oracle.oracle import OracleSim
import rex_gym_b200 as R

def run(task, signal, n, steps, terrain="plane", **kw):
    rng = np.random.default_rng(7)
    env = R.BatchedRexEnv(task=task, num_envs=n, signal_type=signal, terrain_type=terrain, num_fields=4, **kw)
    okw = dict(kw)
    ora = OracleSim(n, task, signal, terrain=terrain, nfields=4, **okw)
    o_g = env.reset(); o_c = ora.reset()
    sg = env.get_state()
    so = [ora.state(i) for i in range(n)]
    print(f"[{task}-{signal}-{terrain}] reset: pos err {max(np.abs(sg['pos'][i]-so[i]['pos']).max() for i in range(n)):.2e} "
          f"q err {max(np.abs(sg['q'][i]-so[i]['q']).max() for i in range(n)):.2e} obs err {np.abs(o_g-o_c).max():.2e}")
    b = R.envs.batched_env.ACTION_BOUND[(task, signal)]
    alive = np.ones(n, bool)
    worst_q = worst_p = 0.0
    for k in range(steps):
        a = rng.uniform(-b, b, size=(n, env.action_dim)).astype(np.float32)
        og, rg, dg, _ = env.step(a)
        oc, rc, dc = ora.step(a)
        sg = env.get_state()
        eq = np.array([np.abs(sg['q'][i] - ora.state(i)['q']).max() for i in range(n)])
        ep = np.array([np.abs(sg['pos'][i] - ora.state(i)['pos']).max() for i in range(n)])
        cm = np.array([ora.env(i).contact_mask & 0x1FF for i in range(n)])
        mism = (cm != sg['contact_mask']) & alive
        worst_q = max(worst_q, eq[alive].max() if alive.any() else 0); worst_p = max(worst_p, ep[alive].max() if alive.any() else 0)
        if k % 100 == 0 or k == steps - 1:
            print(f"  step {k:4d} alive {alive.sum():3d} max|dq| {eq[alive].max() if alive.any() else 0:.2e} max|dpos| {ep[alive].max() if alive.any() else 0:.2e} "
                  f"|dobs| {np.abs(og-oc)[alive].max() if alive.any() else 0:.2e} |drew| {np.abs(rg-rc)[alive].max() if alive.any() else 0:.2e} contact mismatches {mism.sum()} done g/c {dg.sum()}/{dc.sum()}")
        newly = (dg | dc) & alive
        if (dg != dc)[alive].any():
            print(f"  step {k}: done mismatch on {np.nonzero((dg != dc) & alive)[0][:8]}")
        alive &= ~(dg | dc)
        if not alive.any():
            print("  all envs done at", k); break
    print(f"  worst over run: |dq| {worst_q:.3e} |dpos| {worst_p:.3e}; err flags OR = {env.check_errors()}")
    env.close()

if __name__ == "__main__":
    t = time.time()
    run("walk", "ik", 32, 1000, target_position=3.0, backwards=True)
    run("gallop", "ol", 32, 400, target_position=2.0)
    run("gallop", "ik", 32, 300, target_position=2.0)
    run("turn", "ik", 32, 400)
    print("elapsed", time.time() - t)
