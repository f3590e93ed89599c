"""The ctypes stub of INTEGRATION.md section 1, runnable (kept in sync by hand): create / reset / step through the bare C ABI."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ctypes as C, numpy as np, torch
from rex_gym_b200 import _capi                      # RexSimConfig mirrors the struct in include/rexsim.h
from rex_gym_b200.model_tables import pack_model_tables

L = _capi.load()
cfg = _capi.RexSimConfig()
cfg.num_envs, cfg.task, cfg.signal, cfg.terrain = 4096, 0, 0, 0          # walk, ik, plane
cfg.num_motors, cfg.action_repeat, cfg.solver_iterations = 12, 5, 60      # rex_gym_env.py:25,184
cfg.sim_dt_d = 0.005 / 5                                                   # walk_env.py:34-35
cfg.motor_kp = cfg.kp_lo = cfg.kp_hi = 1.0
cfg.motor_kd = cfg.kd_lo = cfg.kd_hi = 0.02
cfg.target_position, cfg.backwards = float("nan"), -1                      # random per reset, like walk_env.py:133-147
cfg.target_orient = cfg.init_orient = float("nan")
cfg.w_distance, cfg.w_energy, cfg.w_drift, cfg.w_shake = 1.0, 0.0005, 2.0, 0.005
cfg.normalize, cfg.max_episode_steps, cfg.auto_reset, cfg.seed = 1, 2000, 0, 1234
cfg.friction, cfg.residual_threshold, cfg.erp_contact, cfg.erp_joint = 0.5, 1e-7, 0.08, 0.2
tables, cfg.toe_npts = pack_model_tables("base"); cfg.toe_margin = -0.00025  # model_tables.TOE_MARGIN
for k in range(5): cfg.pose_values[k] = float("nan")                        # poses task only
cfg.gait_clock_scale = 1.0                                                 # GaitPlanner clock = simulation clock (9 ~ the wall clock of the reference's walk-ik training)
cfg.contact_breaking, cfg.link_damping, cfg.max_coordinate_velocity = 0.00081, 0.04, 100.0   # model_tables
cfg.control_latency = cfg.pd_latency = 0.0                                 # sensor model off (reference default); noise_stdev[5] = 0
sim = C.c_void_p()
assert L.rexsim_create(C.byref(cfg), tables.ctypes.data_as(C.c_void_p), tables.size, C.byref(sim)) == 0

N, A, O = 4096, 2, 4
act = torch.zeros(N, A, device="cuda"); obs = torch.zeros(N, O, device="cuda")
rew = torch.zeros(N, device="cuda");    done = torch.zeros(N, dtype=torch.uint8, device="cuda")
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
L.rexsim_reset(sim, None, 0, C.c_void_p(obs.data_ptr()), stream)           # BatchEnv.reset(None)
L.rexsim_step(sim, C.c_void_p(act.data_ptr()), C.c_void_p(obs.data_ptr()),
              C.c_void_p(rew.data_ptr()), C.c_void_p(done.data_ptr()), stream)   # BatchEnv.step(action)

torch.cuda.synchronize()
print("obs", obs[:2].cpu().numpy(), "reward", rew[:2].cpu().numpy(), "done", int(done.sum()))
assert torch.isfinite(obs).all() and L.rexsim_launch_count(sim) >= 3
