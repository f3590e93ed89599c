"""CPU suite, world_size 2 over gloo: the N>1 path (contiguous env sharding, draws keyed on the global env
id, one all-gather of the packed outputs) reproduces the single-process batch exactly.  The oracle stands
in for the CUDA library here (no GPU in this container); the sharding/collective code under test is the
product's (rex_gym_b200/sharding.py)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

G, STEPS = 8, 12


def _actions():
    return np.random.default_rng(5).uniform(-1, 1, size=(STEPS, G, 2)).astype(np.float32)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import OracleSim
    from rex_gym_b200.sharding import shard_range, all_gather_outputs
    off, n = shard_range(G, rank, world)
    sim = OracleSim(n, "walk", "ik", normalize=True, max_episode_steps=6, settle=False, env_offset=off, seed=99)
    sim.reset()
    acts = _actions()
    rows = []
    for k in range(STEPS):
        o, r, d = sim.step(acts[k, off:off + n])
        if d.any():
            sim.reset(np.nonzero(d)[0])
        oa, ra, da = all_gather_outputs(torch.from_numpy(o), torch.from_numpy(r), torch.from_numpy(d))
        rows.append(np.concatenate([oa.numpy().reshape(-1), ra.numpy(), da.numpy().astype(np.float32)]))
    targets = np.array([sim.env(i).target_position for i in range(n)])
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.stack(rows))
    np.save(os.path.join(out_dir, f"targets{rank}.npy"), targets)
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    np.testing.assert_array_equal(a, b)                       # every rank sees the same gathered batch
    from oracle.oracle import OracleSim
    sim = OracleSim(G, "walk", "ik", normalize=True, max_episode_steps=6, settle=False, seed=99)
    sim.reset()
    acts = _actions()
    for k in range(STEPS):
        o, r, d = sim.step(acts[k])
        if d.any():
            sim.reset(np.nonzero(d)[0])
        ref = np.concatenate([o.reshape(-1), r, d.astype(np.float32)])
        np.testing.assert_array_equal(a[k], ref)              # bit-identical to the unsharded run
    t = np.concatenate([np.load(tmp_path / "targets0.npy"), np.load(tmp_path / "targets1.npy")])
    np.testing.assert_array_equal(t, [sim.env(i).target_position for i in range(G)])   # draws keyed on the global id


def test_shard_range():
    from rex_gym_b200.sharding import shard_range, pack_outputs, unpack_outputs
    assert [shard_range(65536, r, 4) for r in range(4)] == [(0, 16384), (16384, 16384), (32768, 16384), (49152, 16384)]
    with pytest.raises(ValueError):
        shard_range(10, 0, 4)
    o, r, d = torch.rand(5, 4), torch.rand(5), torch.tensor([1, 0, 0, 1, 0], dtype=torch.uint8)
    o2, r2, d2 = unpack_outputs(pack_outputs(o, r, d))
    assert torch.equal(o, o2) and torch.equal(r, r2) and torch.equal(d.bool(), d2)
