#!/usr/bin/env python3
"""Developer tool: step-kernel time against episode age (all envs start together at reset; uniform random actions)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import rex_gym_b200 as R
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
kw = dict(task="walk", signal_type="ik", target_position=2.0, backwards=False)
if len(sys.argv) > 2 and sys.argv[2] == "bw": kw["backwards"] = True
if len(sys.argv) > 2 and sys.argv[2] == "fast": kw["gait_clock_scale"] = 16.0
env = R.BatchedRexEnv(num_envs=n, normalize=True, auto_reset=True, max_episode_steps=2000, **kw)
env.reset()
acts = torch.rand((64, n, env.action_dim), device="cuda") * 2 - 1
ev = [torch.cuda.Event(enable_timing=True) for _ in range(701)]
dones = []
ev[0].record()
for k in range(700):
    _, _, d, _ = env.step(acts[k % 64])
    ev[k + 1].record()
    if k % 50 == 49: dones.append(int(d.sum().item()))
torch.cuda.synchronize()
for c in range(14):
    ms = sum(ev[i].elapsed_time(ev[i + 1]) for i in range(50 * c, 50 * c + 50)) / 50
    print(f"steps {50*c:4d}-{50*c+49:4d}: {ms:.4f} ms/step  {n/ms/1e3:7.1f} M env-steps/s   done at the last step of the chunk: {dones[c]}")
print("flags", env.check_errors())
