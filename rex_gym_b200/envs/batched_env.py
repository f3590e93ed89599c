"""BatchedRexEnv -- host-side mirror of the reference's env/BatchEnv surface over the CUDA library.

Reference interface mirrored (same names, argument meaning, error behaviour):
  * BatchEnv.step / reset / __len__ / __getitem__ / close   rex_gym/agents/tools/batch_env.py:18-115
  * RexWalkEnv / RexReactiveEnv / RexTurnEnv constructor kwargs rex_gym/envs/gym/walk_env.py:31-50,
    gallop_env.py:43-63, turn_env.py:30-49 (+ num_envs, device, and the fused training wrappers)
  * gym ids RexWalk-v0 ... rex_gym/playground/__init__.py:17-57 via make()
One call = one kernel launch for all N environments; there is no CPU fallback.
"""
import ctypes as C
import math

import numpy as np
import torch

from .. import _capi
from ..model_tables import (pack_model_tables, contact_breaking_distance, TOE_MARGIN, LINK_DAMPING,
                            MAX_COORDINATE_VELOCITY)
from ..terrain import make_random_fields
from .spaces import Box

TASKS = {"walk": 0, "gallop": 1, "turn": 2, "standup": 3, "poses": 4}
SIGNALS = {"ik": 0, "ol": 1}
TERRAINS = {"plane": 0, "random": 1}
DEFAULT_URDF_VERSION = "default"
OBSERVATION_EPS = 0.01          # rex_gym/envs/rex_gym_env.py:19
ACTION_BOUND = {("walk", "ik"): 0.4, ("walk", "ol"): 0.01, ("gallop", "ik"): 0.4, ("gallop", "ol"): 0.3,
                ("turn", "ik"): 0.01, ("turn", "ol"): 0.01, ("standup", "ol"): 0.1, ("standup", "ik"): 0.1,
                ("poses", "ik"): 0.1, ("poses", "ol"): 0.1}

ERR_NONFINITE, ERR_JOINT_LIMIT, ERR_BODY_CONTACT, ERR_TILE_MISS, ERR_BAD_INDEX = 1, 2, 4, 8, 16
SENSOR_NOISE_STDDEV = (0.0, 0.0, 0.0, 0.0, 0.0)      # rex_gym/model/rex.py:22, default of rex_gym_env.py:61


class _DevArray(object):
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy, no ownership)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
        self._owner = owner


class _Info(object):
    """Stand-in for BatchEnv's `tuple(infos)`: info[i] == {'action': motor command of env i} of THE STEP THAT RETURNED IT
    (rex_gym_env.py:414).  The command block lives on the device and is overwritten by the next step, so the host copy is taken
    on first access; an access after the batch has stepped again raises instead of handing out the later step's commands."""

    def __init__(self, env):
        self._env, self._cmd, self._at = env, None, env._nsteps

    def __len__(self):
        return len(self._env)

    def materialize(self):
        """Take the host copy now (one device sync + D2H copy of [N][num_motors] floats)."""
        if self._cmd is None:
            if self._env._nsteps != self._at:
                raise RuntimeError("info of step %d read after the batch advanced to step %d: call info.materialize() (or read "
                                   "info[i]) before the next step()" % (self._at, self._env._nsteps))
            self._cmd = self._env.last_command().t().contiguous().cpu().numpy()
        return self

    def __getitem__(self, i):
        return {"action": self.materialize()._cmd[i]}


class _EnvView(object):
    """What BatchEnv.__getitem__ returns: a handle on one environment of the batch."""

    def __init__(self, batch, index):
        self._b, self.index = batch, index
        self.observation_space, self.action_space = batch.observation_space, batch.action_space

    @property
    def env_step_counter(self):
        return int(self._b.get_state()["env_step_counter"][self.index])


class BatchedRexEnv(object):
    metadata = {"render.modes": ["rgb_array"], "video.frames_per_second": 66}

    def __init__(self, task="walk", num_envs=1, device="cuda:0", debug=False, urdf_version=None,
                 control_time_step=None, action_repeat=None, control_latency=0.0, pd_latency=0.0, on_rack=False,
                 motor_kp=1.0, motor_kd=0.02, render=False, num_steps_to_log=2000, env_randomizer=None,
                 log_path=None, target_position=None, backwards=None, target_orient=None, init_orient=None,
                 energy_weight=None, signal_type="ik", terrain_type="plane", terrain_id=None, mark="base",
                 normalize=False, max_episode_steps=0, auto_reset=False, seed=1234,
                 motor_kp_range=None, motor_kd_range=None, num_fields=64, solver_iterations=None, env_offset=0,
                 base_y=None, base_z=None, base_roll=None, base_pitch=None, base_yaw=None, gait_clock_scale=1.0,
                 rebalance_every=8, observation_noise_stdev=SENSOR_NOISE_STDDEV):
        if urdf_version is not None and urdf_version != DEFAULT_URDF_VERSION:
            raise ValueError("%s is not a supported urdf_version." % urdf_version)     # rex_gym_env.py:317-318
        if task not in TASKS or signal_type not in SIGNALS:
            raise ValueError("unknown task/signal_type %r/%r" % (task, signal_type))
        if terrain_type not in TERRAINS:
            raise ValueError("terrain_type %r: only 'plane' and 'random' are built (csv/png assets live in the "
                             "pybullet_data pip package)" % terrain_type)
        if render or on_rack:
            raise ValueError("render / on_rack are GUI debugging modes of the reference; not part of the batched path")
        if control_latency < 0 or pd_latency < 0 or len(observation_noise_stdev) != 5 or min(observation_noise_stdev) < 0:
            raise ValueError("control_latency / pd_latency must be >= 0 and observation_noise_stdev five values >= 0")
        if env_randomizer:
            raise ValueError("env_randomizer hooks are Python callbacks; use motor_kp_range / motor_kd_range")
        if not torch.cuda.is_available():
            raise RuntimeError("rex_gym_b200 needs a CUDA device: there is no CPU fallback")
        self._L = _capi.load()
        self._debug = bool(debug)
        self.task, self.signal_type, self.terrain_type, self.mark = task, signal_type, terrain_type, mark
        self.device = torch.device(device)
        self.num_envs = int(num_envs)
        self.num_motors = 12 if mark == "base" else 18
        rep = action_repeat or (6 if task in ("gallop", "poses") else 5)
        cts = control_time_step or (0.006 if task in ("gallop", "poses") else 0.005)
        self.control_time_step, self._action_repeat = cts, rep
        self._time_step = cts / rep
        if mark not in ("base", "arm"):
            raise ValueError("mark must be 'base' or 'arm'")          # mark_constants.py:1
        tables, toe_npts = pack_model_tables(mark)
        c = _capi.RexSimConfig()
        c.num_envs, c.task, c.signal, c.terrain = self.num_envs, TASKS[task], SIGNALS[signal_type], TERRAINS[terrain_type]
        c.num_motors, c.action_repeat = self.num_motors, rep
        c.solver_iterations = solver_iterations or int(300 / rep)                      # rex_gym_env.py:25,184
        c.sim_dt, c.sim_dt_d = self._time_step, self._time_step
        c.motor_kp, c.motor_kd = motor_kp, motor_kd
        c.kp_lo, c.kp_hi = motor_kp_range or (motor_kp, motor_kp)
        c.kd_lo, c.kd_hi = motor_kd_range or (motor_kd, motor_kd)
        # `if not self._target_position` (walk_env.py:144, gallop_env.py:150): None AND 0 mean "draw one per reset"
        c.target_position = float("nan") if not target_position else target_position
        c.backwards = -1 if backwards is None else int(bool(backwards))
        c.target_orient = float("nan") if target_orient is None else target_orient
        c.init_orient = float("nan") if init_orient is None else init_orient
        c.w_distance, c.w_drift, c.w_shake = 1.0, 2.0, 0.005                            # rex_gym_env.py:56-59
        c.w_energy = energy_weight if energy_weight is not None else (0.005 if task == "gallop" else 0.0005)
        c.normalize, c.max_episode_steps, c.auto_reset, c.seed = int(normalize), int(max_episode_steps), int(auto_reset), seed
        c.gait_clock_scale = float(gait_clock_scale)     # GaitPlanner clock / simulation clock (the reference reads the wall clock)
        for k, v in enumerate((base_y, base_z, base_roll, base_pitch, base_yaw)):      # poses_env.py:49-53 (None = rotate per reset)
            c.pose_values[k] = float("nan") if v is None else float(v)
        # sensor model (rex_gym_env.py:61,70-71 -> Rex(control_latency, pd_latency, observation_noise_stdev), rex.py:726-769)
        c.control_latency, c.pd_latency = float(control_latency), float(pd_latency)
        for k, v in enumerate(observation_noise_stdev):
            c.noise_stdev[k] = float(v)
        self._fields = None
        with torch.cuda.device(self.device):
            if terrain_type == "random":
                self._fields = torch.from_numpy(make_random_fields(num_fields)).to(self.device).contiguous()
                c.nfields, c.fields = num_fields, self._fields.data_ptr()
                c.friction = 0.5 * 0.5          # URDF link default 0.5 x createMultiBody default 0.5
            else:
                c.nfields, c.fields = 0, None
                c.friction = 0.5 * 1.0          # x plane.urdf lateral_friction 1
            c.residual_threshold, c.erp_contact, c.erp_joint = 1e-7, 0.08, 0.2
            c.toe_npts, c.toe_margin = toe_npts, TOE_MARGIN
            c.contact_breaking = contact_breaking_distance(mark)
            c.link_damping, c.max_coordinate_velocity = LINK_DAMPING, MAX_COORDINATE_VELOCITY
            c.env_offset = int(env_offset)
            self._cfg = c
            h = C.c_void_p()
            tb = np.ascontiguousarray(tables)
            _capi.check(self._L.rexsim_create(C.byref(c), tb.ctypes.data, tb.size, C.byref(h)))
        self._h = h
        self.obs_dim = self._L.rexsim_obs_dim(c.task, c.num_motors)
        self.action_dim = self._L.rexsim_action_dim(c.task, c.signal)
        N, O, A = self.num_envs, self.obs_dim, self.action_dim
        dev = self.device
        # device path: obs | reward | done of the last step (torch views; `done` is a bool view of the u8 the kernel writes)
        self._obs = torch.zeros((N, O), dtype=torch.float32, device=dev)
        self._reward = torch.zeros((N,), dtype=torch.float32, device=dev)
        self._done_u8 = torch.zeros((N,), dtype=torch.uint8, device=dev)
        self._done = self._done_u8.view(torch.bool)
        # host path (numpy in / numpy out): ONE pinned block each way, filled by rexsim_step_host in one C call
        nb_obs, nb_rew = N * O * 4, N * 4
        nb_out = int(self._L.rexsim_host_out_bytes(self._h))
        self._h_act = torch.zeros((N, A), dtype=torch.float32).pin_memory()
        self._h_out = torch.zeros((nb_out,), dtype=torch.uint8).pin_memory()
        self._h_act_np = self._h_act.numpy()
        self._h_act_ptr, self._h_out_ptr = self._h_act.data_ptr(), self._h_out.data_ptr()
        ho = self._h_out.numpy()
        self._h_obs = ho[:nb_obs].view(np.float32).reshape(N, O)
        self._h_reward = ho[nb_obs:nb_obs + nb_rew].view(np.float32)
        self._h_done = ho[nb_obs + nb_rew:nb_obs + nb_rew + N].view(np.bool_)
        self._h_err = ho[nb_out - 12:nb_out - 8].view(np.int32)      # int32 OR of the step's error bits (8 scratch bytes follow)
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        p = C.c_void_p()
        _capi.check(self._L.rexsim_error_flags(self._h, C.byref(p)))
        self._err = torch.as_tensor(_DevArray(p.value, (N + 1,), "<i4", self), device=dev)
        _capi.check(self._L.rexsim_last_command(self._h, C.byref(p)))
        self._cmd = torch.as_tensor(_DevArray(p.value, (self.num_motors, N), "<f4", self), device=dev)
        nf, ni = C.c_int32(), C.c_int32()
        _capi.check(self._L.rexsim_state_words(C.byref(c), C.byref(nf), C.byref(ni)))
        pf, pi = C.c_void_p(), C.c_void_p()
        _capi.check(self._L.rexsim_state_buffers(self._h, C.byref(pf), C.byref(pi)))
        self._state_f = torch.as_tensor(_DevArray(pf.value, (nf.value, N), "<f4", self), device=dev)
        self._state_i = torch.as_tensor(_DevArray(pi.value, (ni.value, N), "<i4", self), device=dev)
        pr, nr = C.c_void_p(), C.c_int64()
        _capi.check(self._L.rexsim_history_buffer(self._h, C.byref(pr), C.byref(nr)))          # sensor history ring (None: model off)
        self._ring = torch.as_tensor(_DevArray(pr.value, (nr.value,), "<f4", self), device=dev) if nr.value else None
        # spaces: raw task spaces, or the wrapper-visible ones when the training wrappers are fused in
        b = ACTION_BOUND[(task, signal_type)]
        if normalize:
            self.action_space = Box(-np.inf * np.ones(A, np.float32), np.inf * np.ones(A, np.float32))   # ClipAction wrappers.py:257-260
            self.observation_space = Box(-np.ones(O, np.float32), np.ones(O, np.float32))                # RangeNormalize :204-209
        else:
            hi = np.full(A, b, np.float32)
            self.action_space = Box(hi, -hi) if task == "gallop" else Box(-hi, hi)                       # gallop_env.py:128-130
            ub = np.full(O, 2 * math.pi, np.float32)
            ub[2:4] = 2 * math.pi / self._time_step
            self.observation_space = Box(-(ub + OBSERVATION_EPS), ub + OBSERVATION_EPS)
        # every `rebalance_every` steps the envs are re-grouped over the warps by solver cost (rexsim_rebalance): a scheduling
        # hint only -- results are bit-identical with 0 (off).  It pays off once the batch is several waves deep (+17-19 % at
        # 16 384 / 65 536 envs); a batch that fits one wave (4096 envs = 128 CTAs on 148 SMs) is bound by its slowest env anyway
        self._rebalance_every = int(rebalance_every) if self.num_envs >= 8192 else 0
        self._nsteps = 0
        self._closed = False

    # ---- BatchEnv surface ------------------------------------------------------------------------
    def __len__(self):
        return self.num_envs

    def __getitem__(self, index):
        if not -self.num_envs <= index < self.num_envs:
            raise IndexError(index)
        return _EnvView(self, index % self.num_envs)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _maybe_rebalance(self):
        self._nsteps += 1
        if self._rebalance_every and self._nsteps % self._rebalance_every == 0:
            with torch.cuda.device(self.device):
                _capi.check(self._L.rexsim_rebalance(self._h, self._stream()))

    def step(self, action):
        """BatchEnv.step (batch_env.py:63-90).  numpy in -> numpy out (host buffers: one rexsim_step_host call = H2D copy,
        kernel, D2H copy, wait);  CUDA tensor in -> CUDA tensors out (one asynchronous kernel launch, no host traffic)."""
        N, A = self.num_envs, self.action_dim
        if isinstance(action, torch.Tensor) and action.is_cuda:
            if tuple(action.shape) != (N, A):
                raise ValueError("Invalid action shape %s, expected %s" % (tuple(action.shape), (N, A)))
            if action.dtype != torch.float32 or not action.is_contiguous():
                action = action.to(torch.float32).contiguous()
            if torch.cuda.current_device() != self._dev_index:
                with torch.cuda.device(self.device):
                    rc = self._L.rexsim_step(self._h, action.data_ptr(), self._obs.data_ptr(), self._reward.data_ptr(),
                                             self._done_u8.data_ptr(), self._stream())
            else:
                rc = self._L.rexsim_step(self._h, action.data_ptr(), self._obs.data_ptr(), self._reward.data_ptr(),
                                         self._done_u8.data_ptr(), torch.cuda.current_stream().cuda_stream)
            if rc:
                _capi.check(rc)
            self._maybe_rebalance()
            return self._obs, self._reward, self._done, _Info(self)
        a = np.asarray(action, dtype=np.float32)
        if a.shape != (N, A):
            raise ValueError("Invalid action shape %s, expected %s" % (a.shape, (N, A)))
        if not np.isfinite(a).all():
            raise ValueError("Invalid action: non-finite values")
        np.copyto(self._h_act_np, a)
        if torch.cuda.current_device() != self._dev_index:
            with torch.cuda.device(self.device):
                rc = self._L.rexsim_step_host(self._h, self._h_act_ptr, self._h_out_ptr, self._stream())
        else:
            rc = self._L.rexsim_step_host(self._h, self._h_act_ptr, self._h_out_ptr, torch.cuda.current_stream().cuda_stream)
        if rc:
            _capi.check(rc)
        self._maybe_rebalance()
        if self._h_err[0] & ERR_NONFINITE:        # the aggregate word is per step on this path (rexsim_step_host clears it)
            raise ValueError("Infinite observation encountered.")          # ConvertTo32Bit wrappers.py:522-523,542-543
        return self._h_obs.copy(), self._h_reward.copy(), self._h_done.copy(), _Info(self)

    def step_into(self, action, obs, reward, done_u8):
        """Raw device form of step(): CUDA float32 action [N][A] in; obs [N][O], reward [N] (float32) and done [N] (uint8) are
        written into the caller's contiguous CUDA buffers (e.g. slices of a rollout).  One asynchronous kernel launch."""
        with torch.cuda.device(self.device):
            _capi.check(self._L.rexsim_step(self._h, action.data_ptr(), obs.data_ptr(), reward.data_ptr(), done_u8.data_ptr(), self._stream()))
        self._maybe_rebalance()

    def reset(self, indices=None):
        """BatchEnv.reset (batch_env.py:92-109): observations of the reset environments."""
        with torch.cuda.device(self.device):
            if indices is None:
                k, idx_ptr = self.num_envs, None
            else:
                if isinstance(indices, torch.Tensor):
                    # device indices are not inspected on the host: the kernel skips any index outside [0, N) and raises
                    # REXSIM_FLAG_BAD_INDEX in the aggregate error word (check_errors() -> IndexError); debug=True checks here
                    idx = indices.to(device=self.device, dtype=torch.int32).contiguous()
                    if self._debug and idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= self.num_envs):
                        raise IndexError("reset index out of range")
                else:
                    ia = np.asarray(indices, dtype=np.int64).reshape(-1)
                    if ia.size and (ia.min() < 0 or ia.max() >= self.num_envs):
                        raise IndexError("reset index out of range")
                    idx = torch.from_numpy(ia.astype(np.int32)).to(self.device)
                k, idx_ptr = int(idx.numel()), idx.data_ptr()
            out = torch.zeros((k, self.obs_dim), dtype=torch.float32, device=self.device)
            if k > 0:
                _capi.check(self._L.rexsim_reset(self._h, idx_ptr, k, out.data_ptr(), self._stream()))
            if isinstance(indices, torch.Tensor):
                return out
            return out.cpu().numpy()

    def close(self):
        if not self._closed:
            self._closed = True
            torch.cuda.synchronize(self.device)
            self._L.rexsim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- extras ------------------------------------------------------------------------------------
    def get_state(self):
        """Physical state of every env (pybullet getBasePositionAndOrientation/getBaseVelocity/getJointState)."""
        N, nm = self.num_envs, self.num_motors
        f = torch.zeros((13 + 2 * nm, N), dtype=torch.float32, device=self.device)
        i = torch.zeros((4, N), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _capi.check(self._L.rexsim_get_state(self._h, f.data_ptr(), i.data_ptr(), self._stream()))
        f, i = f.cpu().numpy(), i.cpu().numpy()
        return dict(pos=f[0:3].T.copy(), quat=f[3:7].T.copy(), linvel=f[7:10].T.copy(), angvel=f[10:13].T.copy(),
                    q=f[13:13 + nm].T.copy(), qd=f[13 + nm:13 + 2 * nm].T.copy(),
                    step_counter=i[0].copy(), env_step_counter=i[1].copy(), flags=i[2].copy(), contact_mask=i[3].copy())

    def set_state(self, pos, quat, linvel, angvel, q, qd):
        N, nm = self.num_envs, self.num_motors
        f = np.concatenate([np.asarray(x, np.float32).reshape(N, -1).T for x in (pos, quat, linvel, angvel, q, qd)], axis=0)
        t = torch.from_numpy(np.ascontiguousarray(f)).to(self.device)
        with torch.cuda.device(self.device):
            _capi.check(self._L.rexsim_set_state(self._h, t.data_ptr(), self._stream()))
            torch.cuda.current_stream(self.device).synchronize()

    def last_command(self):
        """[num_motors, N] motor commands of the last step (info['action'], rex_gym_env.py:414)."""
        return self._cmd

    def error_flags(self):
        return self._err[:self.num_envs]

    def check_errors(self):
        """Read AND clear the aggregate error word (every bit raised since the last call; one device sync).  Raises like the
        reference would have for the offending step: ValueError for a non-finite observation (ConvertTo32Bit), IndexError for a
        reset index outside the batch; returns the remaining (informational) bits."""
        e = int(self._err[self.num_envs].item())
        if e:
            with torch.cuda.device(self.device):
                _capi.check(self._L.rexsim_clear_errors(self._h, self._stream()))
        if e & ERR_NONFINITE:
            raise ValueError("Infinite observation encountered.")
        if e & ERR_BAD_INDEX:
            raise IndexError("reset index out of range")
        return e

    def state_dict(self):
        """Exact checkpoint of the whole batch (the reference never checkpoints env state)."""
        sd = {"state_f": self._state_f.clone(), "state_i": self._state_i.clone()}
        if self._ring is not None:
            sd["sensor_history"] = self._ring.clone()
        return sd

    def load_state_dict(self, sd):
        self._state_f.copy_(sd["state_f"])
        self._state_i.copy_(sd["state_i"])
        if self._ring is not None:
            self._ring.copy_(sd["sensor_history"])

    @property
    def launch_count(self):
        return int(self._L.rexsim_launch_count(self._h))

    def seed(self, seed=None):
        return [seed]

    def render(self, mode="rgb_array", close=False):
        return np.array([])


class RexWalkBatchEnv(BatchedRexEnv):
    """rex_gym/envs/gym/walk_env.py:16 RexWalkEnv, batched."""

    def __init__(self, num_envs=1, **kw):
        kw.setdefault("control_time_step", 0.005); kw.setdefault("action_repeat", 5)
        super().__init__(task="walk", num_envs=num_envs, **kw)


class RexGallopBatchEnv(BatchedRexEnv):
    """rex_gym/envs/gym/gallop_env.py:28 RexReactiveEnv, batched."""

    def __init__(self, num_envs=1, **kw):
        kw.setdefault("control_time_step", 0.006); kw.setdefault("action_repeat", 6)
        kw.setdefault("energy_weight", 0.005)
        super().__init__(task="gallop", num_envs=num_envs, **kw)


class RexTurnBatchEnv(BatchedRexEnv):
    """rex_gym/envs/gym/turn_env.py:20 RexTurnEnv, batched."""

    def __init__(self, num_envs=1, **kw):
        kw.setdefault("control_time_step", 0.005); kw.setdefault("action_repeat", 5)
        super().__init__(task="turn", num_envs=num_envs, **kw)


class RexPosesBatchEnv(BatchedRexEnv):
    """rex_gym/envs/gym/poses_env.py:21 RexPosesEnv, batched (no settle at reset; reward 1, never 'fallen')."""

    def __init__(self, num_envs=1, **kw):
        kw.setdefault("control_time_step", 0.006); kw.setdefault("action_repeat", 6)
        super().__init__(task="poses", num_envs=num_envs, **kw)


class RexStandupBatchEnv(BatchedRexEnv):
    """rex_gym/envs/gym/standup_env.py:17 RexStandupEnv, batched (starts from INIT_POSES['rest_position'])."""

    def __init__(self, num_envs=1, **kw):
        kw.setdefault("control_time_step", 0.005); kw.setdefault("action_repeat", 5)
        kw.setdefault("signal_type", "ol")
        super().__init__(task="standup", num_envs=num_envs, **kw)


# gym ids of rex_gym/playground/__init__.py:17-57 -> batched classes
ENV_IDS = {"RexWalk-v0": RexWalkBatchEnv, "RexGalloping-v0": RexGallopBatchEnv, "RexTurn-v0": RexTurnBatchEnv,
           "RexStandup-v0": RexStandupBatchEnv, "RexPoses-v0": RexPosesBatchEnv}


# max_episode_steps of the reference's registrations (playground/__init__.py:17-57): gym.make wraps the env in a TimeLimit
REGISTERED_MAX_EPISODE_STEPS = {"RexWalk-v0": 2500, "RexGalloping-v0": 1000, "RexTurn-v0": 1000, "RexStandup-v0": 400,
                                "RexPoses-v0": 400}


def make(env_id, num_envs=1, **kwargs):
    """gym.make(id, **args) equivalent (rex_gym/playground/trainer.py:47) returning a whole batch; the registered
    max_episode_steps (gym's TimeLimit) is fused into the kernel unless the caller overrides it."""
    if env_id not in ENV_IDS:
        raise ValueError("env id %r not built (have %s)" % (env_id, sorted(ENV_IDS)))
    kwargs.setdefault("max_episode_steps", REGISTERED_MAX_EPISODE_STEPS[env_id])
    return ENV_IDS[env_id](num_envs=num_envs, **kwargs)
