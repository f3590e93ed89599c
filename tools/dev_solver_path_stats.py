"""Dev: which solver path do the sub-steps of each BASELINE workload need?  Oracle statistics over de-synchronised batches:
fraction of env-steps with body-box contacts / joint-limit rows, PGS iteration counts, resets per step."""
import sys, numpy as np
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle.oracle import OracleSim
def run(name, task, sig, steps=600, n=64, **kw):
    o = OracleSim(n, task, sig, max_episode_steps=2000, normalize=True, **kw)
    o.reset()
    rng = np.random.default_rng(0)
    gen = lim = body = 0; iters = []; tot = 0; resets = 0; limrows=[]
    for k in range(steps):
        a = rng.uniform(-1, 1, size=(n, o.A)).astype(np.float32)
        obs, r, d = o.step(a, 8)
        for i in range(n):
            e = o.env(i)
            b = (e.contact_mask & 0b010101011) != 0   # base bit0, upper bits 1,3,5,7
            l = e.limit_rows > 0
            gen += (b or l); lim += l; body += b; tot += 1; iters.append(e.solver_iters); limrows.append(e.limit_rows)
        idx = np.nonzero(d)[0]
        if len(idx): o.reset(idx); resets += len(idx)
    print(f"{name:28s} generic {gen/tot:.3f} (limit {lim/tot:.3f} body {body/tot:.3f}) mean limit rows {np.mean(limrows):.2f} iters mean {np.mean(iters):.1f} p90 {np.percentile(iters,90):.0f} max {np.max(iters)}  resets/step {resets/tot:.4f}")
run("walk-ik scale1", "walk", "ik", target_position=2.0, backwards=False)
run("walk-ik scale4", "walk", "ik", target_position=2.0, backwards=False, gait_clock_scale=4.0)
run("C3 gallop-ol", "gallop", "ol", kp_range=(0.8,1.2), kd_range=(0.01,0.03))
run("C4 turn-ik heightfield", "turn", "ik", terrain="random", nfields=8, steps=400)
run("standup base", "standup", "ol", steps=400)
run("C5 standup arm", "standup", "ol", mark="arm", steps=400)
