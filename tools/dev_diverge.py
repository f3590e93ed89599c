#!/usr/bin/env python3
"""Find the first env/step where GPU and oracle part ways and print what happened around it."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from oracle.oracle import OracleSim
import rex_gym_b200 as R
from rex_gym_b200.envs.batched_env import ACTION_BOUND

def run(task, sig, n, steps, seed, thr=2e-4, **kw):
    env = R.BatchedRexEnv(task=task, num_envs=n, signal_type=sig, seed=seed, **kw)
    okw = dict(kw); kr, dr = okw.pop("motor_kp_range", None), okw.pop("motor_kd_range", None)
    ter = okw.pop("terrain_type", "plane"); nf = okw.pop("num_fields", 0)
    ora = OracleSim(n, task, sig, seed=seed, kp_range=kr, kd_range=dr, terrain=ter, nfields=nf, **okw)
    env.reset(); ora.reset()
    sg = env.get_state()
    print(f"[{task}-{sig}] reset dq {max(np.abs(sg['q'][i]-ora.state(i)['q']).max() for i in range(n)):.2e}")
    rng = np.random.default_rng(11); b = ACTION_BOUND[(task, sig)]
    hist = []
    for k in range(steps):
        a = rng.uniform(-b, b, size=(n, env.action_dim)).astype(np.float32)
        env.step(a); ora.step(a)
        sg = env.get_state()
        eq = np.array([np.abs(sg['q'][i] - ora.state(i)['q']).max() for i in range(n)])
        cm = np.array([ora.env(i).contact_mask & 0x1FF for i in range(n)])
        lim = np.array([ora.env(i).limit_rows for i in range(n)])
        it = np.array([ora.env(i).solver_iters for i in range(n)])
        hist.append((eq, cm, sg['contact_mask'].copy(), lim, it))
        bad = np.nonzero(eq > thr)[0]
        if len(bad):
            i = bad[0]
            print(f"  first divergence: env {i} at step {k}: dq {eq[i]:.2e}; flags {int(env.error_flags()[i])}")
            for kk in range(max(0, k - 8), k + 1):
                h = hist[kk]
                print(f"   step {kk}: dq {h[0][i]:.2e} oracle mask {h[1][i]:09b} gpu mask {h[2][i]:09b} limit rows {h[3][i]} iters {h[4][i]} pos z {ora.env(i).pos[2]:.3f}")
            break
    else:
        print("  no divergence above", thr, "max", max(h[0].max() for h in hist))
    env.close()

if __name__ == "__main__":
    run("gallop", "ol", 32, 120, 3, target_position=2.0, motor_kp_range=(0.8, 1.2), motor_kd_range=(0.01, 0.03))
    run("turn", "ik", 16, 30, 9, terrain_type="random", num_fields=4)
    run("standup", "ol", 16, 60, 3)
