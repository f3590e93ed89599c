#!/usr/bin/env python3
"""Golden for the constant tables of the path, read from the reference's own modules imported from
/root/reference: pose tables (`rex_constants.INIT_POSES`, `ARM_POSES`), motor name order (`mark_constants`), the module constants
of rex.py (init orientation / rack position, overheat shutdown torque and time, sensor noise defaults), motor.py and
kinematics.py / gait_planner.py geometry.  Output: tests/golden/constants_golden.json"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_golden import install_stubs  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "constants_golden.json")


def plain(v):
    if isinstance(v, (list, tuple, np.ndarray)):
        return [plain(x) for x in v]
    if isinstance(v, (np.floating, float)):
        return float(v)
    if isinstance(v, (np.integer, int)):
        return int(v)
    return v


def main():
    install_stubs()
    from rex_gym.model import rex_constants, mark_constants, rex, motor, kinematics, terrain
    from rex_gym.envs import rex_gym_env
    k = kinematics.Kinematics()
    out = {"source": "module constants of rex_gym/model/{rex_constants,mark_constants,rex,motor,kinematics,terrain}.py and envs/rex_gym_env.py",
           "init_poses": {n: plain(v) for n, v in rex_constants.INIT_POSES.items()},
           "arm_poses": {n: plain(v) for n, v in rex_constants.ARM_POSES.items()},
           "motor_names": mark_constants.MARK_DETAILS["motors_names"], "motors_num": mark_constants.MARK_DETAILS["motors_num"],
           "rex": {n: plain(getattr(rex, n)) for n in ("INIT_RACK_POSITION", "INIT_ORIENTATION", "OVERHEAT_SHUTDOWN_TORQUE",
                                                         "OVERHEAT_SHUTDOWN_TIME", "SENSOR_NOISE_STDDEV", "LEG_POSITION")},
           "motor": {n: plain(getattr(motor, n)) for n in ("VOLTAGE_CLIPPING", "OBSERVED_TORQUE_LIMIT", "MOTOR_VOLTAGE", "MOTOR_RESISTANCE",
                                                             "MOTOR_TORQUE_CONSTANT", "MOTOR_VISCOUS_DAMPING", "MOTOR_SPEED_LIMIT")},
           "kinematics": {n: plain(getattr(k, n)) for n in ("_l", "_w", "_hip", "_leg", "_foot", "y_dist", "x_dist", "height")
                          if hasattr(k, n)},
           "env": {"OBSERVATION_EPS": rex_gym_env.OBSERVATION_EPS, "NUM_SIMULATION_ITERATION_STEPS": rex_gym_env.NUM_SIMULATION_ITERATION_STEPS,
                   "MOTOR_ANGLE_OBSERVATION_INDEX": rex_gym_env.MOTOR_ANGLE_OBSERVATION_INDEX},
           "robot_init_position": terrain.ROBOT_INIT_POSITION}
    mm = motor.MotorModel(12)
    out["motor"]["current_table"] = plain(mm._current_table)
    out["motor"]["torque_table"] = plain(mm._torque_table)
    with open(OUT, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    print(out["kinematics"], out["rex"])


if __name__ == "__main__":
    main()
