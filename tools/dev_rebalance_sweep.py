#!/usr/bin/env python3
"""Developer tool: step time of a de-synchronised 65 536-env walk-ik batch against the re-grouping period."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import rex_gym_b200 as R
from bench import stagger_episodes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
for every in (0, 1, 2, 4, 8, 16, 32):
    env = R.BatchedRexEnv(task="walk", num_envs=n, signal_type="ik", normalize=True, auto_reset=True, max_episode_steps=2000,
                          target_position=2.0, backwards=False, rebalance_every=every)
    env.reset()
    acts = torch.rand((32, n, 2), device="cuda") * 2 - 1
    stagger_episodes(env, acts)
    for k in range(40): env.step(acts[k % 32])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for k in range(300): env.step(acts[k % 32])
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 300
    print(f"rebalance_every={every:2d}: {ms:.4f} ms/step  {n/ms/1e3:.1f} M env-steps/s", flush=True)
    env.close()
