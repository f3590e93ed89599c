#!/usr/bin/env python3
"""Developer tool: where the time of one host-buffer BatchedRexEnv.step goes (4096 envs)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import rex_gym_b200 as R
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = R.BatchedRexEnv(task="walk", num_envs=n, signal_type="ik", normalize=True, auto_reset=True, max_episode_steps=2000, target_position=2.0, backwards=False)
env.reset()
a = np.random.default_rng(0).uniform(-1, 1, (n, env.action_dim)).astype(np.float32)
for _ in range(50): env.step(a)
K = 500
t0 = time.perf_counter()
for _ in range(K): env.step(a)
full = (time.perf_counter() - t0) / K
L, h = env._L, env._h
st = torch.cuda.current_stream().cuda_stream
t0 = time.perf_counter()
for _ in range(K): L.rexsim_step_host(h, env._h_act_ptr, env._h_out_ptr, st)
ccall = (time.perf_counter() - t0) / K
act = torch.from_numpy(a).cuda()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(K): env.step(act)
torch.cuda.synchronize()
dev = (time.perf_counter() - t0) / K
t0 = time.perf_counter()
for _ in range(K):
    b = np.asarray(a, dtype=np.float32); ok = np.isfinite(b).all(); np.copyto(env._h_act_np, b)
    o = env._h_obs.copy(); r = env._h_reward.copy(); d = env._h_done.copy()
py = (time.perf_counter() - t0) / K
print(f"n={n}: full step {full*1e6:.1f} us | C call (kernel + flags copy + sync) {ccall*1e6:.1f} us | numpy in/out work {py*1e6:.1f} us | device path (async, amortised) {dev*1e6:.1f} us")
