#!/usr/bin/env python3
"""Developer tool: SM clock / power while the graph-captured rollout runs flat out (is the sustained rollout power-capped?)."""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import rex_gym_b200 as R
from rex_gym_b200.agents import ForwardGaussianPolicy, Rollout
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = R.BatchedRexEnv(task="walk", num_envs=n, signal_type="ik", normalize=True, auto_reset=True, max_episode_steps=2000, target_position=2.0, backwards=False)
net = ForwardGaussianPolicy(env.obs_dim, env.action_dim)
ro = Rollout(env, net, 32, training=True)
ro.collect(); torch.cuda.synchronize()
samples, stop = [], threading.Event()
def smi():
    while not stop.is_set():
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,temperature.gpu", "--format=csv,noheader,nounits"], stdout=subprocess.PIPE, text=True)
        samples.append(r.stdout.strip()); stop.wait(0.25)
th = threading.Thread(target=smi); th.start()
for reps in (10, 100, 300):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): ro.collect()
    b.record(); torch.cuda.synchronize()
    print(f"{reps} windows: {a.elapsed_time(b) / reps / 32:.4f} ms per control step")
stop.set(); th.join()
print("clocks.sm, power W, sw_power_cap, hw_slowdown, temp:", samples[::3])
