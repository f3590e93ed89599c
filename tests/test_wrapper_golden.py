"""SURVEY row a15 pinned to the reference's own wrapper classes.

tests/golden/wrapper_golden.json.gz (tools/gen_wrapper_golden.py) was produced by running
rex_gym.agents.tools.wrappers.{LimitDuration, RangeNormalize, ClipAction, ConvertTo32Bit}, stacked as
rex_gym/playground/trainer.py:49-52 stacks them, around a scripted inner env whose action / observation Boxes come from the
task envs' own constructor lines.  The oracle's restatement of that arithmetic (`rexo_wrap_action`, `rexo_wrap_observation`,
the LimitDuration counter in `step_env`) is checked against it; the CUDA path is checked against the oracle with
`normalize=True` throughout tests/test_gpu_parity.py, and the host-side space tables against the same fixture here."""
import gzip
import json
import math
import os

import numpy as np
import pytest

from oracle.oracle import OracleSim

G = json.load(gzip.open(os.path.join(os.path.dirname(__file__), "golden", "wrapper_golden.json.gz"), "rt"))
CASES = [(c["task"], c["signal"]) for c in G["cases"]]


def _case(task, signal):
    return next(c for c in G["cases"] if c["task"] == task and c["signal"] == signal)


def test_fixture_covers_every_task_and_the_reference_wrapper_order():
    assert G["wrapper_order"] == ["LimitDuration", "RangeNormalize", "ClipAction", "ConvertTo32Bit"]
    assert set(CASES) == {("walk", "ik"), ("walk", "ol"), ("gallop", "ik"), ("gallop", "ol"), ("turn", "ik"), ("turn", "ol"),
                          ("standup", "ol"), ("poses", "ik")}


@pytest.mark.parametrize("task,signal", CASES)
def test_action_and_observation_maps_match_the_reference_wrappers(task, signal):
    c = _case(task, signal)
    s = OracleSim(1, task, signal, normalize=True, settle=2)
    assert s.cfg.sim_dt == c["sim_dt"]                              # control_time_step / action_repeat of the task env
    acts, inner = np.array(c["policy_actions"]), np.array(c["inner_actions"])
    assert s.A == acts.shape[1] and s.O == len(c["observ_low"])
    for a, want in zip(acts, inner):                                # ClipAction then _denormalize_action: what the task env receives
        np.testing.assert_allclose(s.wrap_action(a), want, rtol=0, atol=1e-15)
    lo, hi = np.array(c["action_low"]), np.array(c["action_high"])
    assert np.all(np.abs(inner) <= np.abs(lo) + 1e-15)              # never outside the task Box, whatever the policy emits
    if task == "gallop":
        assert np.all(lo > hi)                                      # the inverted Box (gallop_env.py:128-130) ...
        np.testing.assert_allclose(s.wrap_action(np.ones(s.A)), hi, atol=1e-15)      # ... maps +1 to its `high` = -b
    # observations: every wrapped observation of the scripted episode, float32 as ConvertTo32Bit leaves them
    raw, seen = np.array(c["raw_observs"]), np.array(c["observs"], np.float32)
    order = [0] + list(range(1, len(seen)))                         # reset observation first, then one per step
    for k, want in zip(order, seen):
        got = s.wrap_observation(raw[k])
        assert got.dtype == np.float32
        np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(s.wrap_observation(c["observ_high"]), np.ones(s.O, np.float32))
    np.testing.assert_array_equal(s.wrap_observation(c["observ_low"]), -np.ones(s.O, np.float32))
    raw_s = OracleSim(1, task, signal, normalize=False, settle=2)   # wrappers off: identity on both sides
    np.testing.assert_array_equal(raw_s.wrap_action(acts[0]), acts[0])
    np.testing.assert_array_equal(raw_s.wrap_observation(raw[1]), raw[1].astype(np.float32))


def test_limit_duration_counts_like_the_reference():
    """LimitDuration (wrappers.py:266-291): `done` on the step whose count reaches `duration`, the count restarts on reset,
    an env-side `done` passes through unchanged.  The fixture's inner env ends its first episode with the 5th step; after the
    reset the 9th step is cut by the limit (duration 9) and stepping on raises."""
    c = _case("walk", "ik")
    assert c["dones"] == [False] * 4 + [True] + [False] * 8 + [True] and c["reset_before_step"] == 5 and c["duration"] == 9
    assert c["step_after_limit"][1] == "Must reset environment."
    s = OracleSim(2, "walk", "ik", normalize=True, settle=2, max_episode_steps=c["duration"], target_position=2.0, backwards=False)
    s.reset()
    dones = []
    for t in range(14):
        if t == 5:
            s.reset(np.array([0, 1], np.int32))
        _, _, d = s.step(np.zeros((2, 2), np.float32))
        dones.append(bool(d[0]))
    assert dones == [False] * 13 + [True]                           # 5 steps, reset, then exactly `duration` more
    s.reset(np.array([1], np.int32))                                # per-env counters: env 1 restarts, env 0 is past its limit
    _, _, d = s.step(np.zeros((2, 2), np.float32))
    assert d[0] and not d[1]


def test_host_side_space_tables_match_the_reference_boxes():
    """BatchedRexEnv's raw and wrapper-visible spaces are built from ACTION_BOUND / OBSERVATION_EPS (batched_env.py); the same
    numbers the reference's constructors produce."""
    from rex_gym_b200.envs import batched_env as B
    for c in G["cases"]:
        b = B.ACTION_BOUND[(c["task"], c["signal"])]
        lo, hi = np.array(c["action_low"]), np.array(c["action_high"])
        want_lo = b if c["task"] == "gallop" else -b
        np.testing.assert_array_equal(lo, np.full(lo.shape, want_lo)); np.testing.assert_array_equal(hi, -lo)
        ub = np.full(len(c["observ_high"]), 2 * math.pi); ub[2:4] = 2 * math.pi / c["sim_dt"]
        np.testing.assert_allclose(c["observ_high"], ub + B.OBSERVATION_EPS, rtol=1e-15)
        np.testing.assert_allclose(c["observ_low"], -(ub + B.OBSERVATION_EPS), rtol=1e-15)
        wl, wh, ol, oh = [np.array(x) for x in c["wrapped_spaces"]]
        assert np.all(np.isneginf(wl)) and np.all(np.isposinf(wh))  # ClipAction shows an unbounded Box (wrappers.py:257-260)
        np.testing.assert_array_equal(ol, -1.0); np.testing.assert_array_equal(oh, 1.0)
    assert c["nonfinite_error"] == "Infinite observation encountered."   # the message BatchedRexEnv raises (tests/test_gpu_parity.py)
