#!/usr/bin/env python3
"""Golden vectors for the sensor-history path: the reference's own `Rex.ReceiveObservation`,
`Rex._GetDelayedObservation`, `Rex._GetPDObservation` and `Rex._GetControlObservation` (rex_gym/model/rex.py:726-761),
imported from /root/reference and run UNMODIFIED on a Rex object created with object.__new__ whose `GetTrueObservation` is the
scripted closed-form state of tests/golden/script.py (no pybullet needed: these methods only touch the history deque).

Output: tests/golden/sensor_golden.json.gz -- for every control/pd latency pair and every sub-step k of the script:
the control observation and the PD observation the reference returns after ReceiveObservation number k.
"""
import collections
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "tests", "golden"))
from gen_golden import install_stubs, euler_to_quat  # noqa: E402
from script import scripted_state  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "sensor_golden.json.gz")
DT, NM, STEPS = 0.001, 12, 130
LATENCIES = [(0.0, 0.0), (0.001, 0.0003), (0.0032, 0.002), (0.02, 0.005), (0.099, 0.0999), (0.2, 0.15)]


def true_obs(k):
    pos, rpy, angvel, q, qd, tau = scripted_state(k, 0.3, 0.1, 0.2, 1, DT)
    return list(q) + list(qd) + list(tau) + list(euler_to_quat(rpy)) + list(angvel)


def main():
    install_stubs()
    from rex_gym.model import rex as rexmod
    out = {"dt": DT, "num_motors": NM, "steps": STEPS, "script": "scripted_state(k, 0.3, 0.1, 0.2, 1, dt): q, qd, tau, quat(rpy), angvel",
           "source": "rex_gym/model/rex.py:726-761 run unmodified", "cases": []}
    for cl, pl in LATENCIES:
        r = object.__new__(rexmod.Rex)
        r._observation_history = collections.deque(maxlen=100)
        r.time_step, r.num_motors = DT, NM
        r._control_latency, r._pd_latency = cl, pl
        k = [0]
        r.GetTrueObservation = lambda: true_obs(k[0])
        ctrl, pdq = [], []
        for i in range(STEPS):
            k[0] = i
            r.ReceiveObservation()
            ctrl.append([float(x) for x in np.asarray(r._control_observation)])
            q, qd = r._GetPDObservation()
            pdq.append([float(x) for x in np.concatenate([q, qd])])
        out["cases"].append({"control_latency": cl, "pd_latency": pl, "control_observation": ctrl, "pd_observation": pdq})
    with gzip.open(OUT, "wt") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
