#!/usr/bin/env python3
"""Golden vectors for SURVEY row a15 (the training wrappers) and the task envs' action / observation spaces, produced by
running the reference's own code from /root/reference (read-only):

  rex_gym.agents.tools.wrappers.{LimitDuration, RangeNormalize, ClipAction, ConvertTo32Bit}   (wrappers.py:183-291,461-544)
      stacked in the order rex_gym/playground/trainer.py:49-52 stacks them, around a scripted inner env
  RexWalkEnv / RexReactiveEnv / RexTurnEnv / RexStandupEnv / RexPosesEnv .__init__              (the lines that build
      `action_space`: walk_env.py:104-114, gallop_env.py:119-130, turn_env.py:100-110, standup_env.py:99-101,
      poses_env.py:115-117) and ._get_observation_{upper,lower}_bound + OBSERVATION_EPS (rex_gym_env.py:19,277-282)

Shims (recorded in the fixture): stub modules gym / pybullet / pybullet_data / tensorflow (absent here); RexGymEnv.__init__
(which opens a pybullet client and loads the URDF) is replaced by a recorder that keeps the keyword arguments, so each task
env's OWN constructor lines run unmodified after it.

Output: tests/golden/wrapper_golden.json.gz (committed; the GPU box has no /root/reference).
"""
import gzip
import json
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import install_stubs, FakeRex  # noqa: E402

OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "wrapper_golden.json.gz")


def install_tf_stub():
    tf = types.ModuleType("tensorflow")
    log = types.SimpleNamespace(info=lambda *a, **k: None, warning=lambda *a, **k: None, error=lambda *a, **k: None)
    tf.logging = log
    tf.compat = types.SimpleNamespace(v1=types.SimpleNamespace(logging=log))
    sys.modules["tensorflow"] = tf


class ScriptedEnv:
    """Inner env: records the action it is handed, returns scripted float64 observations / rewards / dones."""

    def __init__(self, action_space, observation_space, observs, rewards, dones):
        self.action_space, self.observation_space = action_space, observation_space
        self._o, self._r, self._d = observs, rewards, dones
        self.k = 0
        self.seen = []

    def reset(self):
        self.k = 0
        return np.array(self._o[0], np.float64)

    def step(self, action):
        self.seen.append(np.asarray(action, np.float64).tolist())
        self.k += 1
        return np.array(self._o[self.k], np.float64), float(self._r[self.k]), bool(self._d[self.k]), {}


def main():
    install_stubs()
    install_tf_stub()
    import rex_gym.envs.rex_gym_env as rge
    from rex_gym.agents.tools import wrappers

    def recorder_init(self, **kw):                       # stands in for RexGymEnv.__init__ (pybullet client, URDF, reset)
        self._kw = kw
        self._signal_type = kw.get("signal_type")
        self._time_step = kw["control_time_step"] / kw["action_repeat"]     # rex_gym_env.py:170-173
        self.mark = kw.get("mark", "base")
        self._on_rack = kw.get("on_rack", False)
        self.num_motors = 12 if self.mark == "base" else 18
        self.rex = FakeRex(self._time_step, self.num_motors)     # scripted getters: only _get_observation_dimension needs them
        self.rex.num_motors = self.num_motors
    rge.RexGymEnv.__init__ = recorder_init
    from rex_gym.envs.gym import walk_env, gallop_env, turn_env, standup_env, poses_env
    classes = {"walk": walk_env.RexWalkEnv, "gallop": gallop_env.RexReactiveEnv, "turn": turn_env.RexTurnEnv,
               "standup": standup_env.RexStandupEnv, "poses": poses_env.RexPosesEnv}
    gold = {"reference": "nicrusso7/rex-gym @ /root/reference",
            "shims": ["stub gym/pybullet/pybullet_data/tensorflow", "RexGymEnv.__init__ -> keyword recorder"],
            "wrapper_order": ["LimitDuration", "RangeNormalize", "ClipAction", "ConvertTo32Bit"], "cases": []}
    rng = np.random.default_rng(20260923)
    for task, cls in classes.items():
        for signal in (("ik", "ol") if task in ("walk", "gallop", "turn") else ("ol",) if task == "standup" else ("ik",)):
            kw = dict(signal_type=signal) if task != "standup" else {}
            env = cls.__new__(cls)
            try:
                env.__init__(**kw)
            except AttributeError as e:                       # poses: the lines after `action_space` read state the real base
                assert task == "poses" and hasattr(env, "action_space"), e      # constructor would have set (poses_env.py:124)
            asp = env.action_space
            hi = env._get_observation_upper_bound() + rge.OBSERVATION_EPS
            lo = env._get_observation_lower_bound() - rge.OBSERVATION_EPS
            osp = sys.modules["gym.spaces"].Box(lo, hi)
            A, O = asp.shape[0], osp.shape[0]
            T, duration = 16, 9
            # raw observations inside and at the edge of the box; rewards; the inner env ends its first episode with the 5th step
            obs = rng.uniform(-1, 1, (T + 1, O)) * hi * rng.choice([1e-3, 0.1, 1.0], (T + 1, 1))
            obs[3] = hi; obs[4] = lo
            rew = rng.normal(0, 1, T + 1)
            don = np.zeros(T + 1, bool); don[5] = True
            inner = ScriptedEnv(asp, osp, obs, rew, don)
            env_w = wrappers.LimitDuration(inner, duration)
            env_w = wrappers.RangeNormalize(env_w)
            env_w = wrappers.ClipAction(env_w)
            env_w = wrappers.ConvertTo32Bit(env_w)
            # policy-side actions: inside, on and far outside [-1, 1]
            acts = rng.uniform(-1, 1, (T, A)) * rng.choice([0.5, 1.0, 1.7, 7.0], (T, 1))
            acts[2] = 1.0; acts[3] = -1.0
            o0 = env_w.reset()
            out_o, out_r, out_d = [o0], [], []
            after_limit = None
            for t in range(T):
                if t == 5:                                   # the 5th step ended the episode (inner done): reset, LimitDuration restarts
                    o = env_w.reset(); inner.k = 5
                    assert np.array_equal(o, out_o[0])
                try:
                    o, r, d, _ = env_w.step(acts[t])
                except RuntimeError as e:                    # stepping past LimitDuration's end without a reset (wrappers.py:275-276)
                    after_limit = (t, str(e))
                    break
                out_o.append(o); out_r.append(r); out_d.append(bool(d))
            assert all(o.dtype == np.float32 for o in out_o) and all(r.dtype == np.float32 for r in out_r)
            space_w = (env_w.action_space.low.tolist(), env_w.action_space.high.tolist(),
                       env_w.observation_space.low.tolist(), env_w.observation_space.high.tolist())
            bad = None
            try:                                             # ConvertTo32Bit on a non-finite observation (wrappers.py:522-523)
                inner2 = ScriptedEnv(asp, osp, [obs[0], np.full(O, np.nan)], [0, 0], [False, False])
                w2 = wrappers.ConvertTo32Bit(wrappers.ClipAction(wrappers.RangeNormalize(wrappers.LimitDuration(inner2, 5))))
                w2.reset(); w2.step(acts[0])
            except ValueError as e:
                bad = str(e)
            gold["cases"].append(dict(task=task, signal=signal, sim_dt=env._time_step, action_low=asp.low.tolist(), action_high=asp.high.tolist(),
                                      observ_low=lo.tolist(), observ_high=hi.tolist(), wrapped_spaces=[[float(x) for x in s] for s in space_w],
                                      policy_actions=acts.tolist(), inner_actions=inner.seen, raw_observs=obs.tolist(),
                                      observs=[np.asarray(o, np.float64).tolist() for o in out_o], rewards=[float(r) for r in out_r],
                                      raw_rewards=rew.tolist(), dones=out_d, duration=duration, reset_before_step=5,
                                      step_after_limit=after_limit, nonfinite_error=bad))
            print(task, signal, "A", A, "O", O, "action box", asp.low[0], asp.high[0], "dones", out_d)
    with gzip.open(OUT, "wt") as f:
        json.dump(gold, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
