#!/usr/bin/env python3
"""Generate golden vectors for the CONTROLLER half of the hot path by importing the reference's
own Python modules from /root/reference (read-only) and running them here.

What runs for real (reference code, unmodified):
  rex_gym.model.motor.MotorModel.convert_to_torque          (model/motor.py:76-143)
  rex_gym.model.kinematics.Kinematics.solve                 (model/kinematics.py:104-142)
  rex_gym.model.gait_planner.GaitPlanner.loop               (model/gait_planner.py:96-134)
  RexWalkEnv / RexReactiveEnv / RexTurnEnv / RexStandupEnv
      ._transform_action_to_motor_command, ._reward, ._termination, ._get_observation
      (envs/gym/*.py, envs/rex_gym_env.py:490-542)
with three shims, all recorded in the fixture header:
  * numpy.math = math                  (np.math was removed in NumPy 2; gait_planner.py:24)
  * gait_planner.time.time -> sim clock (the reference uses wall-clock, gait_planner.py:108-110;
                                         the deterministic replacement is t = step_counter*dt)
  * stub modules gym / pybullet / pybullet_data (absent here); the env objects are created with
    object.__new__ and given a scripted fake `rex` (base pose, joint state), so only the pure-Python
    signal / reward / termination bodies execute.  pybullet's three pure quaternion helpers are
    restated in the stub (they are third-party, not reference code).

Output: tests/golden/controller_golden.json.gz  (committed; the GPU box has no /root/reference).
"""
import json
import math
import os
import random
import sys
import types

import numpy as np

REF = "/root/reference"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden"))
from script import scripted_state  # noqa: E402
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "controller_golden.json.gz")


# ---------------------------------------------------------------------------------------------
# shims
# ---------------------------------------------------------------------------------------------
def quat_to_euler(q):
    x, y, z, w = q
    sarg = -2 * (x * z - w * y)
    if sarg <= -0.99999:
        return [0, -0.5 * math.pi, 2 * math.atan2(x, -y)]
    if sarg >= 0.99999:
        return [0, 0.5 * math.pi, 2 * math.atan2(-x, y)]
    return [math.atan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z), math.asin(sarg),
            math.atan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z)]


def euler_to_quat(rpy):
    hr, hp, hy = rpy[0] / 2, rpy[1] / 2, rpy[2] / 2
    cr, sr, cp, sp, cy, sy = math.cos(hr), math.sin(hr), math.cos(hp), math.sin(hp), math.cos(hy), math.sin(hy)
    return [sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy]


def quat_to_mat(q):
    x, y, z, w = q
    d = x * x + y * y + z * z + w * w
    s = 2.0 / d
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz, xx, xy, xz, yy, yz, zz = w * xs, w * ys, w * zs, x * xs, x * ys, x * zs, y * ys, y * zs, z * zs
    return [1 - (yy + zz), xy - wz, xz + wy, xy + wz, 1 - (xx + zz), yz - wx, xz - wy, yz + wx, 1 - (xx + yy)]


def install_stubs():
    np.math = math
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")

    class Box:
        def __init__(self, low, high, **kw):
            self.low, self.high = np.asarray(low), np.asarray(high)
            self.shape = self.low.shape
    spaces.Box = Box
    gym.Env = object
    gym.spaces = spaces
    seeding.np_random = lambda seed=None: (np.random.RandomState(seed), seed)
    utils.seeding = seeding
    gym.utils = utils
    pb = types.ModuleType("pybullet")
    pb.getMatrixFromQuaternion = quat_to_mat
    pb.getEulerFromQuaternion = quat_to_euler
    pb.getQuaternionFromEuler = euler_to_quat
    pb.GUI, pb.DIRECT = 1, 2
    pbd = types.ModuleType("pybullet_data")
    pbd.getDataPath = lambda: ""
    for name, mod in (("gym", gym), ("gym.spaces", spaces), ("gym.utils", utils), ("gym.utils.seeding", seeding),
                      ("pybullet", pb), ("pybullet_data", pbd)):
        sys.modules[name] = mod
    sys.path.insert(0, REF)
    return pb


class FakeRex:
    """Scripted stand-in for rex_gym.model.rex.Rex: only the getters the signal/reward code reads."""

    def __init__(self, dt, nm=12):
        self.dt, self.nm = dt, nm
        self.step_counter = 0
        self.pos = [0.0, 0.0, 0.2]
        self.quat = [0.0, 0.0, 0.0, 1.0]
        self.angvel = [0.0, 0.0, 0.0]
        self.q = np.zeros(nm)
        self.qd = np.zeros(nm)
        self.tau = np.zeros(nm)
        self.initial_pose = None

    def GetTimeSinceReset(self):
        return self.step_counter * self.dt

    def GetBasePosition(self):
        return tuple(self.pos)

    def GetBaseOrientation(self):
        return euler_to_quat(quat_to_euler(self.quat))      # rex.py:530-537 round trip

    def GetTrueBaseOrientation(self):
        return list(self.quat)

    def GetBaseRollPitchYaw(self):
        return np.array(quat_to_euler(self.quat))

    GetTrueBaseRollPitchYaw = GetBaseRollPitchYaw

    def GetBaseRollPitchYawRate(self):
        return np.array(self.angvel)

    GetTrueBaseRollPitchYawRate = GetBaseRollPitchYawRate

    def GetMotorTorques(self):
        return np.array(self.tau)

    def GetMotorVelocities(self):
        return np.array(self.qd)

    def GetMotorAngles(self):
        from rex_gym.model.rex import MapToMinusPiToPi
        return MapToMinusPiToPi(np.array(self.q))


def rnd_quat(rng, scale):
    rpy = rng.uniform(-scale, scale, 3)
    return euler_to_quat(rpy)


def main():
    pb = install_stubs()
    from rex_gym.model import motor, kinematics, gait_planner, rex_constants
    gold = {"shims": ["numpy.math=math", "gait_planner.time.time=sim clock", "stub gym/pybullet/pybullet_data"],
            "reference": "nicrusso7/rex-gym @ /root/reference"}
    rng = np.random.default_rng(20260922)

    # ---- motor model ------------------------------------------------------------------------
    mm = motor.MotorModel(12, kp=1.0, kd=0.02)
    cases = []
    for k in range(64):
        q = rng.uniform(-2, 2, 12)
        cmd = q + rng.normal(0, [0.02, 0.2, 1.0][k % 3], 12)
        qd = rng.normal(0, [1, 10, 60][k % 3], 12)
        kp = np.full(12, rng.uniform(0.5, 2.0))
        kd = np.full(12, rng.uniform(0.0, 0.05))
        ta, to = mm.convert_to_torque(cmd, q, qd, qd, kp, kd)
        cases.append(dict(cmd=cmd.tolist(), q=q.tolist(), qd=qd.tolist(), kp=kp.tolist(), kd=kd.tolist(),
                          tau_act=np.asarray(ta).tolist(), tau_obs=np.asarray(to).tolist()))
    stand = rex_constants.INIT_POSES["stand"]
    ta, to = mm.convert_to_torque(stand, stand + np.linspace(-.5, .5, 12), np.linspace(-20, 20, 12), np.linspace(-20, 20, 12))
    gold["motor"] = cases
    gold["motor_survey"] = dict(tau_act=np.asarray(ta).tolist(), tau_obs=np.asarray(to).tolist())

    # ---- IK -----------------------------------------------------------------------------------
    kin = kinematics.Kinematics()
    frames0 = np.asarray(kin._frames).copy()
    cases = []
    for k in range(64):
        rpy = rng.uniform(-0.4, 0.4, 3) if k % 4 else np.zeros(3)
        pos = rng.uniform(-0.03, 0.03, 3)
        fr = frames0 + rng.uniform(-0.05, 0.05, (4, 3)) * (1 if k % 5 else 4)   # some out-of-reach -> domain clamp
        out = kinematics.Kinematics().solve(rpy.copy(), pos.copy(), np.asmatrix(fr))
        cases.append(dict(rpy=rpy.tolist(), pos=pos.tolist(), frames=fr.tolist(),
                          angles=[np.asarray(a).tolist() for a in out[:4]]))
    gold["ik"] = cases

    # ---- gait planner sequences (time patched to a sim clock) -----------------------------------
    clock = {"t": 0.0}
    gait_planner.time.time = lambda: clock["t"]
    seqs = []
    for mode, dt, nsteps in (("walk", 0.005, 280), ("gallop", 0.006, 160), ("walk", 0.005, 320)):
        gp = gait_planner.GaitPlanner(mode)
        steps = []
        vary = len(seqs) == 2
        for k in range(nsteps):
            clock["t"] = k * dt
            v = 0.6 * min(1.0, k * dt) if mode == "walk" else 1.3 * min(1.0, k * dt)
            ang, w_rot, T, d = 0.0, 0.0, (0.65 if mode == "walk" else 0.3), 1.0
            if vary:   # turn-like: rotation + period jitter + step angle, negative speed sometimes
                v, ang, w_rot, T = 0.02, rng.uniform(-30, 30), rng.uniform(-0.6, 0.6), 0.75 + rng.uniform(-0.01, 0.01)
                d = -1.0 if k % 7 == 0 else 1.0
            fr = gp.loop(v, ang, w_rot, T, d)
            steps.append(dict(t=clock["t"], v=v, angle=ang, w_rot=w_rot, T=T, direction=d,
                              frames=np.asarray(fr).tolist(), phi=float(gp._phi), alpha=float(gp._alpha)))
        seqs.append(dict(mode=mode, steps=steps))
    gold["gait"] = seqs
    gold["gait_points"] = dict(
        swing=[dict(phi=p, v=v, angle=a, direction=d,
                    xyz=[float(x) for x in gait_planner.GaitPlanner("walk").calculate_bezier_swing(p, v, a, d)])
               for p, v, a, d in ((0.5, 1.0, 0.0, 1.0), (0.1, -0.3, 20.0, -1.0), (0.93, 0.6, -75.0, 1.0), (0.0, 0.6, 0., 1.), (1.0, 0.6, 0., 1.))],
        stance=[dict(phi=p, v=v, angle=a, xyz=[float(x) for x in gait_planner.GaitPlanner.calculate_stance(p, v, a)])
                for p, v, a in ((0.25, 1.0, 0.0), (0.8, -0.4, 33.0), (0.0, 0.6, 200.0))])

    # ---- env-level signal / reward / termination / observation ---------------------------------
    from rex_gym.envs.gym import walk_env, gallop_env, turn_env, standup_env
    from rex_gym.envs import rex_gym_env

    def make_env(cls, dt, signal, **attrs):
        env = object.__new__(cls)
        env._pybullet_client = pb
        env.mark = "base"
        env._signal_type = signal
        env._time_step = dt
        env._is_render = False
        env._is_debug = False
        env._base_x, env._base_y, env._base_z = 0.01, 0.0, 0.0
        env._base_roll = env._base_pitch = env._base_yaw = 0.0
        env.step_length = env.step_rotation = env.step_angle = env.step_period = None
        env._objectives = []
        env._objective_weights = [1.0, 0.0005, 2.0, 0.005]
        env._forward_reward_cap = float("inf")
        env.env_goal_reached = False
        env.goal_reached = False
        env._stay_still = False
        env.is_terminating = False
        env._kinematics = kinematics.Kinematics()
        env.rex = FakeRex(dt)
        env.rex.initial_pose = rex_constants.INIT_POSES["stand"]
        env.init_pose = rex_constants.INIT_POSES["stand_ol" if signal == "ol" else "stand"]
        for k, v in attrs.items():
            setattr(env, k, v)
        return env

    def run_script(env, act_dim, bound, nsteps, repeat, xspeed, yawspeed=0.0, yaw0=0.0, tilt_at=None):
        """Drive the reference's own step() body (minus physics) along a scripted base trajectory."""
        steps = []
        rex = env.rex
        for k in range(nsteps):
            action = rng.uniform(-bound, bound, act_dim)
            t_pre = rex.GetTimeSinceReset()
            cmd = np.asarray(env._transform_action_to_motor_command(action.copy())).astype(float)
            rex.step_counter += repeat
            pos, rpy, angvel, q, qd, tau = scripted_state(k, xspeed, yawspeed, yaw0, repeat, rex.dt, tilt_at)
            rex.pos, rex.quat, rex.angvel = pos, euler_to_quat(rpy), angvel
            rex.q, rex.qd, rex.tau = np.array(q), np.array(qd), np.array(tau)
            reward = float(env._reward())
            done = bool(env._termination())
            obs = np.asarray(env._get_observation()).astype(float)
            steps.append(dict(t=t_pre, action=action.tolist(), cmd=cmd.tolist(), reward=reward, done=done, obs=obs.tolist(),
                              flags=[int(bool(getattr(env, "goal_reached", False))), int(bool(getattr(env, "_stay_still", False))),
                                     int(bool(env.env_goal_reached))]))
        return steps

    import builtins
    real_print = builtins.print
    builtins.print = lambda *a, **k: None      # the reference prints "FALLING DOWN!" etc.
    envs = []
    try:
        for signal, backwards, adim, bound in (("ik", False, 2, 0.4), ("ik", True, 2, 0.4), ("ol", False, 8, 0.01)):
            clock_env = {}
            env = make_env(walk_env.RexWalkEnv, 0.001, signal, _target_position=0.3, _backwards=backwards, backwards=backwards)
            env._gait_planner = gait_planner.GaitPlanner("walk")
            gait_planner.time.time = (lambda e: (lambda: e.rex.GetTimeSinceReset()))(env)
            xs = 0.5 if backwards else -0.5
            steps = run_script(env, adim, bound, 230, 5, xs, tilt_at=220)
            envs.append(dict(task="walk", signal=signal, backwards=backwards, target_position=0.3, repeat=5, dt=0.001,
                             script=dict(xspeed=xs, yawspeed=0.0, yaw0=0.0, tilt_at=220), steps=steps))
        for signal, adim, bound in (("ik", 2, 0.4), ("ol", 4, 0.3)):
            env = make_env(gallop_env.RexReactiveEnv, 0.001, signal, _target_position=0.4, _backwards=None,
                           _use_angle_in_observation=True)
            env._objective_weights = [1.0, 0.005, 2.0, 0.005]
            env._gait_planner = gait_planner.GaitPlanner("gallop")
            gait_planner.time.time = (lambda e: (lambda: e.rex.GetTimeSinceReset()))(env)
            steps = run_script(env, adim, bound, 230, 6, -0.5, tilt_at=220)
            envs.append(dict(task="gallop", signal=signal, target_position=0.4, repeat=6, dt=0.001,
                             script=dict(xspeed=-0.5, yawspeed=0.0, yaw0=0.0, tilt_at=220), steps=steps))
        for signal, init_o, targ_o in (("ik", 0.5, 1.6), ("ik", 2.0, 0.4), ("ol", 5.5, 1.0)):
            env = make_env(turn_env.RexTurnEnv, 0.001, signal, _target_orient=targ_o, _init_orient=init_o)
            env._gait_planner = gait_planner.GaitPlanner("walk")
            env.clockwise = env._solve_direction()
            gait_planner.time.time = (lambda e: (lambda: e.rex.GetTimeSinceReset()))(env)
            env.rex.quat = euler_to_quat([0, 0, init_o])
            yawspeed = (-1.0 if env.clockwise else 1.0) * 1.6
            steps = run_script(env, 2, 0.01, 420, 5, 0.001, yawspeed=yawspeed, yaw0=init_o)
            envs.append(dict(task="turn", signal=signal, init_orient=init_o, target_orient=targ_o, clockwise=bool(env.clockwise),
                             repeat=5, dt=0.001, script=dict(xspeed=0.001, yawspeed=yawspeed, yaw0=init_o, tilt_at=None), steps=steps))
        env = make_env(standup_env.RexStandupEnv, 0.001, "ol")
        steps = run_script(env, 1, 0.1, 60, 5, 0.0, tilt_at=50)
        envs.append(dict(task="standup", signal="ol", repeat=5, dt=0.001, script=dict(xspeed=0.0, yawspeed=0.0, yaw0=0.0, tilt_at=50), steps=steps))
    finally:
        builtins.print = real_print
    gold["envs"] = envs

    import gzip
    with gzip.GzipFile(OUT, "wb", mtime=0) as f:      # deterministic bytes
        f.write(json.dumps(gold).encode())
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


def poses_golden(pb):
    """RexPosesEnv (envs/gym/poses_env.py): reset's pose rotation + _signal / _reward / _termination / _get_observation,
    written to tests/golden/poses_golden.json.gz (a separate fixture so the first one keeps its bytes)."""
    import collections
    import gzip
    from rex_gym.model import rex_constants
    from rex_gym.envs.gym import poses_env
    rng = np.random.default_rng(20260923)
    out = {"reference": "nicrusso7/rex-gym @ /root/reference", "shims": ["stub gym/pybullet/pybullet_data", "scripted fake rex"],
           "envs": [], "rotation": []}

    def make_env(**pose_args):
        env = object.__new__(poses_env.RexPosesEnv)
        env._pybullet_client = pb
        env.mark = "base"
        env._is_render = False
        env.manual_control = False
        env.load_ui = True
        env._time_step = 0.001
        for k in ("base_y", "base_z", "base_roll", "base_pitch", "base_yaw"):
            setattr(env, "_" + k, pose_args.get(k))
        env._queue = collections.deque(["base_y", "base_z", "roll", "pitch", "yaw"])         # rex_gym_env.py:259-267
        env._ranges = {"base_x": (-0.02, 0.02, 0.01), "base_y": (-0.007, 0.007, 0), "base_z": (-0.048, 0.021, 0),
                       "roll": (-np.pi / 4, np.pi / 4, 0), "pitch": (-np.pi / 4, np.pi / 4, 0), "yaw": (-np.pi / 4, np.pi / 4, 0)}
        env.env_goal_reached = False
        env.rex = FakeRex(0.001)
        env.rex.initial_pose = rex_constants.INIT_POSES["stand"]
        return env

    def reset_tail(env):
        """The body of RexPosesEnv.reset after super().reset() (poses_env.py:148-165), run verbatim on the object."""
        if env._base_y is not None or env._base_z is not None or env._base_roll is not None \
                or env._base_pitch is not None or env._base_yaw is not None:
            env.fill_next_pose_and_target()
        else:
            env.next_pose = env._queue.popleft()
            env._queue.append(env.next_pose)
            env.target_value = random.uniform(env._ranges[env.next_pose][0], env._ranges[env.next_pose][1])
        env.values = env._ranges.copy()

    # deque rotation over successive resets: construction calls reset() once, the user's resets follow
    env = make_env()
    for k in range(12):
        reset_tail(env)
        out["rotation"].append(dict(reset_index=k, next_pose=env.next_pose,
                                    in_range=bool(env._ranges[env.next_pose][0] <= env.target_value <= env._ranges[env.next_pose][1])))
    cases = [dict(base_y=0.005, base_z=0.0, base_roll=0.0, base_pitch=0.0, base_yaw=0.0),
             dict(base_y=0.0, base_z=-0.03, base_roll=0.0, base_pitch=0.0, base_yaw=0.0),
             dict(base_y=0.0, base_z=0.0, base_roll=0.5, base_pitch=0.0, base_yaw=0.0),
             dict(base_y=0.0, base_z=0.0, base_roll=0.0, base_pitch=-0.6, base_yaw=0.0),
             dict(base_y=0.0, base_z=0.0, base_roll=0.0, base_pitch=0.0, base_yaw=0.7),
             dict(base_y=0.0, base_z=0.0, base_roll=0.0, base_pitch=0.0, base_yaw=0.0)]
    names = ["base_y", "base_z", "roll", "pitch", "yaw"]
    for pa in cases:
        env = make_env(**pa)
        reset_tail(env)
        steps = []
        rex = env.rex
        for k in range(200):
            action = rng.uniform(-0.1, 0.1, 1)
            t_pre = rex.GetTimeSinceReset()
            cmd = np.asarray(poses_env.RexPosesEnv._convert_from_leg_model(env._signal(t_pre, action.copy()))).astype(float)
            rex.step_counter += 6
            pos, rpy, angvel, q, qd, tau = scripted_state(k, 0.0, 0.0, 0.0, 6, rex.dt, None)
            rex.pos, rex.quat, rex.angvel = pos, euler_to_quat(rpy), angvel
            steps.append(dict(t=t_pre, action=action.tolist(), cmd=cmd.tolist(), reward=float(env._reward()),
                              done=bool(env.is_fallen()), obs=np.asarray(env._get_observation()).astype(float).tolist()))
        out["envs"].append(dict(args=pa, next_pose=names.index(env.next_pose), target_value=float(env.target_value),
                                repeat=6, dt=0.001, steps=steps))
    path = os.path.join(os.path.dirname(OUT), "poses_golden.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(out).encode())
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    if "--poses-only" in sys.argv:
        poses_golden(install_stubs())
    else:
        main()
        poses_golden(sys.modules["pybullet"])
