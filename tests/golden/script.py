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
