#!/usr/bin/env python3
"""bench.py -- env-steps/sec of the fused step kernel (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W        (N>1: launched under torch.distributed.run)
  python bench.py --impl reference ...                 (CPU arm: the fp64 oracle port on the host cores)

A "step" is one RexGymEnv.step for every environment of the batch (one kernel launch); the workload is
BASELINE.json configs[1]: 4096 envs/GPU, walk-ik, flat terrain, fused ABA+IK+motor kernel, synthetic
uniform random actions, training wrappers (ClipAction/RangeNormalize/LimitDuration) and auto-reset fused.
`value` = device-timed (CUDA events, inputs resident in HBM); `e2e` = through BatchedRexEnv.step with host
(pinned) action buffers in and obs/reward/done back out every step.  Envs shard across GPUs with no data
path collective (weak scaling); one all-gather of the packed outputs is timed separately and reported.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 4096
WORKLOAD = dict(task="walk", signal_type="ik", terrain_type="plane", target_position=2.0, backwards=False,
                normalize=True, max_episode_steps=2000, auto_reset=True)
# algorithmic bytes per env-step (recomputed for the state layout actually built):
# 43 float + 14 int state words read once and written once, + actions, obs, reward, done, last command
B_ALG = 2 * (43 + 14) * 4 + 2 * 4 + 4 * 4 + 4 + 1 + 12 * 4


def clocks_sampler(stop, out, index):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                               stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 6:
                out.append(f)
        except Exception:
            pass
        stop.wait(0.2)


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    mx = max(int(s[1]) for s in samples if s[1].isdigit())
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons}


def stagger_episodes(env, acts, span=256, groups=32):
    """Untimed: spread the episode ages of the batch uniformly over `span` control steps (reset 1/groups of the envs, step
    span/groups times, repeat).  After a common reset every env is in the same phase of the gait ramp and the step kernel's
    cost follows that phase (0.42-0.84 ms per step at 65 536 envs); an auto-resetting training batch is de-synchronised, so
    the timed window should be too -- whatever --steps the caller picks."""
    import torch
    n = env.num_envs
    perm = torch.randperm(n, device=env.device, generator=torch.Generator(device=env.device).manual_seed(7))
    per = (n + groups - 1) // groups
    k = 0
    for g in range(groups):
        idx = perm[g * per:(g + 1) * per].to(torch.int32)
        if idx.numel():
            env.reset(idx)
        for _ in range(span // groups):
            env.step(acts[k % acts.shape[0]]); k += 1


def run_reference(args):
    """CPU arm: the reference's CPU path.  pybullet is not installable here (no network, not in the wheelhouse),
    so this times the fp64 oracle port of the same path (oracle/, kind='port') with all host threads, on a
    bounded sample of the workload (256 envs per step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.oracle import OracleSim
    try:
        cores = len(os.sched_getaffinity(0))        # cores this process may actually use (cgroup/affinity aware)
    except AttributeError:
        cores = os.cpu_count() or 1
    try:                                            # a cgroup CPU quota below the affinity count would oversubscribe
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            cores = min(cores, max(1, int(int(q) / int(p))))
    except Exception:
        pass
    cores = max(1, min(cores, 64))
    os.environ.setdefault("OMP_WAIT_POLICY", "passive")
    n = 32 * cores                                  # 32 envs per thread per step keeps the fork/join overhead small
    sim = OracleSim(n, "walk", "ik", target_position=2.0, backwards=False, normalize=True, max_episode_steps=2000)
    sim.reset()
    rng = np.random.default_rng(1234)
    acts = rng.uniform(-1, 1, size=(args.steps + args.warmup, n, sim.A)).astype(np.float32)
    for k in range(args.warmup):
        _, _, d = sim.step(acts[k], nthreads=cores)
        if d.any():
            sim.reset(np.nonzero(d)[0])
    t0 = time.perf_counter()
    for k in range(args.warmup, args.warmup + args.steps):
        _, _, d = sim.step(acts[k], nthreads=cores)
        if d.any():
            sim.reset(np.nonzero(d)[0])
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    cb = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
          "sample": f"{n} envs x {args.steps} steps of the walk-ik flat workload, OpenMP over envs"}
    print(json.dumps({"impl": "reference", "metric": "env-steps/sec", "value": v, "unit": "env-steps/s",
                      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": f"walk-ik flat, {n}-env sample per step, fp64 CPU port of the path on {cores} threads"},
                      "cpu_baseline": cb,
                      "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def cpu_baseline_leg():
    from oracle.oracle import OracleSim
    n, steps = 64, 200
    sim = OracleSim(n, "walk", "ik", target_position=2.0, backwards=False, normalize=True, max_episode_steps=2000)
    sim.reset()
    rng = np.random.default_rng(1234)
    acts = rng.uniform(-1, 1, size=(steps, n, sim.A)).astype(np.float32)
    t0 = time.perf_counter()
    for k in range(steps):
        _, _, d = sim.step(acts[k], nthreads=1)
        if d.any():
            sim.reset(np.nonzero(d)[0])
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": "env-steps/s", "cores": 1, "kind": "port",
            "sample": f"{n} envs x {steps} steps of the same walk-ik workload, single thread, fp64 oracle port"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import rex_gym_b200 as R
    n = args.envs_per_gpu
    W = max(args.warmup, 3)
    K = args.steps
    env = R.BatchedRexEnv(num_envs=n, device=f"cuda:{local}", seed=1234, env_offset=rank * n, **WORKLOAD)
    env.reset()
    A, O = env.action_dim, env.obs_dim
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    acts = torch.rand((W + K, n, A), device=dev, generator=gen) * 2 - 1          # resident in HBM
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)        # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-timed arm -----------------------------------------------------------------------
    stagger_episodes(env, acts)
    for k in range(W):
        env.step(acts[k])
    barrier()
    stop, samples = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, local), daemon=True); th.start()
    l0 = env.launch_count
    kern_ms, done_count = 0.0, 0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for k in range(K):
        flush.zero_()                                   # L2 flush between timed iterations (not timed)
        ev[k][0].record()
        _, _, d, _ = env.step(acts[W + k])
        ev[k][1].record()
    barrier()
    kern_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = env.launch_count - l0
    t_local = torch.tensor([kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    total_ms = float(t_local.item())
    value = world * n * K / (total_ms * 1e-3)

    # ---- end-to-end arm: host buffers through the public API, copies inside the timed region ------------
    acts_h = acts.cpu().numpy()
    for k in range(3):
        env.step(acts_h[k])
    barrier()
    t0 = time.perf_counter()
    for k in range(K):
        o, r, d, _ = env.step(acts_h[W + k])
    barrier()
    e2e_s = time.perf_counter() - t0
    t_local = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    e2e_value = world * n * K / float(t_local.item())
    stop.set(); th.join(timeout=2)

    # ---- optional observation all-gather (only needed when the learner wants the full batch everywhere) ----
    ag_us = None
    if world > 1:
        packed = torch.zeros((n, O + 2), device=dev)
        outb = torch.zeros((world * n, O + 2), device=dev)
        for _ in range(5):
            dist.all_gather_into_tensor(outb, packed)
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            dist.all_gather_into_tensor(outb, packed)
        b.record(); torch.cuda.synchronize(dev)
        ag_us = a.elapsed_time(b) / 20 * 1e3

    err = env.check_errors()
    env.close()
    # ---- the north-star headline size (65 536 envs per GPU), same workload, device-timed, reported as extras: the nominal
    #      simulation-clock gait, and the gait clock of the reference's own training runs (wall clock ~16x) ----
    extra = None
    if n != 65536:
        nb = 65536
        extra = {}
        for name, kw in (("sim_clock", {}), ("training_clock_x16", {"gait_clock_scale": 16.0})):
            envb = R.BatchedRexEnv(num_envs=nb, device=f"cuda:{local}", seed=1234 + rank, env_offset=rank * nb, **WORKLOAD, **kw)
            envb.reset()
            Kb = min(K, 100)
            actsb = torch.rand((16, nb, A), device=dev, generator=gen) * 2 - 1
            stagger_episodes(envb, actsb)
            for k in range(W):
                envb.step(actsb[k % 16])
            barrier()
            evb = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(Kb)]
            for k in range(Kb):
                flush.zero_()
                evb[k][0].record(); envb.step(actsb[k % 16]); evb[k][1].record()
            barrier()
            tb = torch.tensor([sum(a.elapsed_time(b) for a, b in evb)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tb, op=dist.ReduceOp.MAX)
            extra[name] = {"envs_per_gpu": nb, "steps": Kb, "ms_per_step": float(tb.item()) / Kb,
                           "value": world * nb * Kb / (float(tb.item()) * 1e-3), "unit": "env-steps/s", "error_flags_or": envb.check_errors()}
            envb.close()
        extra.update(extra.pop("sim_clock"))          # the nominal-gait numbers stay at the top level of north_star_size
    # ---- the all-device rollout (policy inference + env step + filter update per control step,
    #      one CUDA graph per 32-step window), same workload, reported as an extra ----------------------------------------
    rollout = None
    try:
        from rex_gym_b200.agents import ForwardGaussianPolicy, Rollout
        envr = R.BatchedRexEnv(num_envs=n, device=f"cuda:{local}", seed=1234, env_offset=rank * n, **WORKLOAD)
        net = ForwardGaussianPolicy(envr.obs_dim, envr.action_dim, device=f"cuda:{local}")
        ro = Rollout(envr, net, 32, seed=1234, training=True, use_graph=True)
        perm = torch.randperm(n, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
        for g in range(8):                  # de-synchronise the episodes: 8 groups, 32 control steps apart (untimed)
            idx = perm[g * n // 8:(g + 1) * n // 8]
            ro._cur[idx] = envr.reset(idx.to(torch.int32))
            ro.collect()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            ro.collect()
        b.record(); barrier()
        tr = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        rollout = {"what": "ForwardGaussianPolicy(200-100) perform + env step + filter update per control step, CUDA graph of 32 steps",
                   "ms_per_control_step": float(tr.item()) / 320, "value": world * n * 320 / (float(tr.item()) * 1e-3), "unit": "env-steps/s",
                   "kernels_per_control_step": 3}
        envr.close(); net.close()
    except Exception as e:  # the extra must never take the headline down with it
        rollout = {"error": str(e)}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        per_launch_s = total_ms * 1e-3 / K
        achieved = n * B_ALG / per_launch_s / 1e9
        line = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{n} envs/GPU walk-ik flat terrain, fused ABA+IK+motor kernel (BASELINE configs[1])",
                       "envs_per_gpu": n, "global_envs": world * n, "parallelism": f"env-sharded x{world}, no data-path collective",
                       "wrappers": "ClipAction+RangeNormalize+LimitDuration(2000)+auto-reset fused", "l2": "flushed between timed steps (256 MiB memset)",
                       "episode_phase": "ages staggered uniformly over 256 control steps before the warm-up (de-synchronised batch, as under auto-reset)",
                       "obs_allgather_us": ag_us, "error_flags_or": err, "north_star_size": extra, "rollout_with_policy": rollout},
            "clocks": summarize_clocks(samples),
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": n * A * 4,
                    "d2h_bytes_per_step": n * (O * 4 + 4 + 1) + 4},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one step_kernel launch at this workload size, from the
                         # ncu --set full capture summarised in profiles/r01b_step_kernel_4096_ncu_raw.txt (the 1 MB state stays in
                         # the 126 MB L2 across launches: reads are first touches of the flushed lines, writes never reach DRAM)
                         "traffic": 987648 if n == 4096 else None, "of": "measured" if peaks else "fallback",
                         "note": "state fits L2; the kernel is issue/latency bound (one partial wave: 0.43 waves/SM, issue slots 24 % busy, "
                                 "14.7 M warp instructions per launch), see DESIGN.md section 5 and profiles/"},
        }
        try:
            line["cpu_baseline"] = cpu_baseline_leg()
        except Exception as e:  # the checker library is test infrastructure; report rather than die
            line["cpu_baseline"] = {"value": None, "unit": "env-steps/s", "cores": 1, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
