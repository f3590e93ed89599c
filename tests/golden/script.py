"""Closed-form scripted "physics" shared by tools/gen_golden.py (which drives the reference's own
signal/reward code with it) and tests/test_controller_golden.py (which drives the oracle with it),
so the fixture only has to store actions and the reference's outputs."""
import math


def scripted_state(k, xspeed, yawspeed, yaw0, repeat, dt, tilt_at=None):
    """State after control step k (0-based): base pose, twist, joint state, observed torques."""
    t = (k + 1) * repeat * dt
    pos = [xspeed * t, 0.02 * math.sin(0.05 * k), 0.2]
    yaw = yaw0 + yawspeed * t
    tilt = 0.05 * math.sin(0.1 * k)
    if tilt_at is not None and k >= tilt_at:
        tilt = 0.7
    rpy = [tilt, 0.5 * tilt, yaw]
    angvel = [0.5 * math.sin(1.3 * k + a) for a in range(3)]
    q = [4.0 * math.sin(0.37 * k + i) for i in range(12)]
    qd = [3.0 * math.cos(0.21 * k + 2 * i) for i in range(12)]
    tau = [5.7 * math.sin(0.11 * k + 3 * i) for i in range(12)]
    return pos, rpy, angvel, q, qd, tau


OVERHEAT_SUBSTEPS, OVERHEAT_REPEAT = 2300, 5


def overheat_joint_state(k):
    """Joint state at sub-step k of the motor-protection script (tools/gen_apply_action_golden.py): small, slow motion so the
    position error decides which side of the 2.45 N m shutdown torque a motor sits on."""
    q = [0.02 * math.sin(0.0126 * k + i) + 0.1 * (i - 6) for i in range(12)]
    qd = [0.02 * 12.6 * math.cos(0.0126 * k + i) for i in range(12)]
    return q, qd


def overheat_command(c):
    """Motor command of control step c: joint angle at the step's first sub-step + a scripted error.  Motor 0 is held 0.5 rad
    off for good; motor 1 gets exactly one control step of relief after 1000 hot sub-steps; motor 2 is relieved right after its
    1001st; motor 3 hovers around the threshold; motor 4 pushes the other way; the rest follow closely."""
    q, _ = overheat_joint_state(c * OVERHEAT_REPEAT)
    err = [0.0] * 12
    err[0] = 0.5
    err[1] = 0.0 if c == 200 else 0.5
    err[2] = 0.5 if c <= 200 else 0.0
    err[3] = 0.215 + 0.05 * math.sin(0.05 * c)
    err[4] = -0.6
    for i in range(5, 12):
        err[i] = 0.05 * math.sin(0.3 * c + i)
    return [a + b for a, b in zip(q, err)]
