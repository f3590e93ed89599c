#!/usr/bin/env python3
"""Workload for ncu captures (profiles/): a few rollout iterations (perform -> step -> experience) at a given batch size.
Usage under ncu:  ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 5 -c 1 -o out python tools/prof_one.py 65536"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import rex_gym_b200 as R  # noqa: E402
from rex_gym_b200.agents import ForwardGaussianPolicy, Rollout  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
task = sys.argv[2] if len(sys.argv) > 2 else "walk"
env = R.BatchedRexEnv(task=task, num_envs=n, signal_type="ik", normalize=True, auto_reset=True, max_episode_steps=2000,
                      **(dict(target_position=2.0, backwards=False) if task == "walk" else {}))
net = ForwardGaussianPolicy(env.obs_dim, env.action_dim)
ro = Rollout(env, net, 12, training=True, use_graph=False)
ro.collect()
torch.cuda.synchronize()
print("done", env.launch_count, net.launch_count)
