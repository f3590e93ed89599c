#!/usr/bin/env python3
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from oracle.oracle import OracleSim
import rex_gym_b200 as R
n, steps, window = 8, 1000, 25
kw = dict(target_position=3.0, backwards=True)
env = R.BatchedRexEnv(task="walk", num_envs=n, **kw); ora = OracleSim(n, "walk", "ik", **kw)
env.reset(); ora.reset()
rng = np.random.default_rng(5)
printed = 0
inwin = np.zeros(n, bool)
for k in range(steps):
    if k % window == 0:
        st = [ora.state(i) for i in range(n)]
        S = {key: np.stack([s[key] for s in st]) for key in st[0]}
        env.set_state(S["pos"], S["quat"], S["linvel"], S["angvel"], S["q"], S["qd"])
        inwin[:] = False
    a = rng.uniform(-0.4, 0.4, size=(n, 2)).astype(np.float32)
    env.step(a); ora.step(a)
    sg = env.get_state()
    si = env._state_i.cpu().numpy()
    for i in range(n):
        dq = np.abs(sg["q"][i] - ora.state(i)["q"])
        if dq.max() > 1e-3 and not inwin[i]:
            inwin[i] = True
            if printed < 4:
                printed += 1
                j = int(dq.argmax()); e = ora.env(i)
                ken = (si[2, i] >> 8) & 0xFFF
                oen = sum((1 if e.enabled[m] else 0) << m for m in range(12))
                kc = [(si[9 + m // 3, i] >> (10 * (m % 3))) & 1023 for m in range(12)]
                oc = [min(e.overheat[m], 1023) for m in range(12)]
                cmdk = env.last_command().cpu().numpy()[:, i]
                print(f"   q k/o {sg['q'][i][j]:.6f}/{ora.state(i)['q'][j]:.6f} qd k/o {sg['qd'][i][j]:.4f}/{ora.state(i)['qd'][j]:.4f} cmd k/o {cmdk[j]:.6f}/{e.cmd[j]:.6f} limit_rows {e.limit_rows} stepctr {e.step_counter} flags {si[2,i] & 0xFF:08b} goal {e.goal_reached} still {e.stay_still} phi {e.gp_phi:.4f} dq vec {np.round(dq,5)}")
                print(f"step {k} env {i} joint {j} dq {dq.max():.2e} enabled k/o {ken:012b}/{oen:012b} counters k {kc} o {oc} masks {int(sg['contact_mask'][i]):09b}/{e.contact_mask & 0x1FF:09b} iters {e.solver_iters}")
print("done")
