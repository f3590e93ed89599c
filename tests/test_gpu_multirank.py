"""GPU, 2 ranks (skipped on a single-GPU box): an environment-sharded rollout -- one process per GPU, NCCL -- reproduces the
single-GPU rollout of the whole batch: reset draws and sampling noise key on the GLOBAL env id, and the streaming
normaliser sees the whole batch through the one all-reduce of this path (rexagent_experience_partial / _finalize)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
N, T = 1024, 6
KW = dict(task="walk", signal_type="ik", normalize=True, auto_reset=True, max_episode_steps=2000, seed=5)


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import rex_gym_b200 as R
    from rex_gym_b200.agents import ForwardGaussianPolicy, Rollout
    n = N // world
    env = R.BatchedRexEnv(num_envs=n, device=f"cuda:{rank}", env_offset=rank * n, **KW)
    net = ForwardGaussianPolicy(env.obs_dim, env.action_dim, device=f"cuda:{rank}", seed=3)
    b = Rollout(env, net, T, seed=17, training=True).collect()
    f = net.get_filters()
    got = {k: b[k].float().cpu().numpy() for k in ("observ", "action", "reward")}
    got["count"], got["mean"], got["var_sum"] = f["observ_count"], f["observ_mean"], f["observ_var_sum"]
    out[rank] = got
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_rollout_reproduces_the_single_gpu_rollout():
    import torch.multiprocessing as mp
    import rex_gym_b200 as R
    from rex_gym_b200.agents import ForwardGaussianPolicy, Rollout
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    env = R.BatchedRexEnv(num_envs=N, device="cuda:0", **KW)
    net = ForwardGaussianPolicy(env.obs_dim, env.action_dim, seed=3)
    b = Rollout(env, net, T, seed=17, training=True, use_graph=False).collect()
    f = net.get_filters()
    for k in ("observ", "action", "reward"):
        whole = b[k].float().cpu().numpy()
        parts = np.concatenate([out[0][k], out[1][k]], axis=1)
        np.testing.assert_allclose(parts, whole, atol=2e-4, err_msg=k)      # sums reduced in a different order: 1e-7 seeds, 6 steps
    np.testing.assert_array_equal(parts[0], whole[0])                       # reset observations + first actions: bit-identical
    for r in range(2):
        assert out[r]["count"] == f["observ_count"] == N * T
        np.testing.assert_allclose(out[r]["mean"], f["observ_mean"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(out[r]["var_sum"], f["observ_var_sum"], rtol=1e-4, atol=1e-6)
    np.testing.assert_array_equal(out[0]["var_sum"], out[1]["var_sum"])     # both ranks hold the same filter state
