#!/usr/bin/env python3
import sys, os, time
t00 = time.time()
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import torch
print("import torch", time.time() - t00, flush=True)
import rex_gym_b200 as R
from rex_gym_b200 import _capi, build
print("needs_build", build.needs_build(), flush=True)
t = time.time(); _capi.load(); print("load lib", time.time() - t, flush=True)
for n in [32, 4096, 65536]:
    t = time.time()
    env = R.BatchedRexEnv(num_envs=n, task="walk", target_position=2.0, backwards=False, normalize=True, auto_reset=True, max_episode_steps=2000)
    torch.cuda.synchronize(); print(n, "create", time.time() - t, flush=True)
    t = time.time(); env.reset(); torch.cuda.synchronize(); print(n, "reset", time.time() - t, flush=True)
    acts = torch.rand((60, n, 2), device="cuda") * 2 - 1
    for k in range(5):
        t = time.time(); env.step(acts[k]); torch.cuda.synchronize(); print(n, "step(dev)", k, time.time() - t, flush=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(5, 55): env.step(acts[k])
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 50
    print(n, "steady ms/step", ms, "env-steps/s", n / ms * 1e3, flush=True)
    ah = acts.cpu().numpy()
    t = time.time()
    for k in range(20): env.step(ah[k])
    print(n, "numpy path ms/step", (time.time() - t) / 20 * 1e3, flush=True)
    t = time.time(); s = env.get_state(); print(n, "get_state", time.time() - t, "err", env.check_errors(), flush=True)
    env.close()
