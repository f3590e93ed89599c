"""Which gait clock does the reference's IK walk actually run on?  Evidence from the reference's own README animation.

`GaitPlanner.loop` takes its phase from the WALL clock (rex_gym/model/gait_planner.py:108-110: time.time()), while the
stride ramp and the brake after the goal run on the SIMULATION clock (walk_env.py:228-244, t = step_counter * dt).  The
deterministic replacement used everywhere in this repo is `gait clock = simulation clock x gait_clock_scale`; SURVEY.md fixed
scale = 1 for the bench workload.  In THIS model the forward trot is marginal at scale 1 (it tips onto the swinging front foot,
0.16 m, ~240 control steps) and walks to its 2 m target for every scale from 1.5 to 5.  Round 1's verdict asked for scale 1 to
walk because "with render=True the reference sleeps to real time, so the README playback runs at about the sim clock".  The
recording itself says otherwise (tests/golden/readme_gif_series.json, measured by tools/gif_measurements.py from
images/walk_ik.gif, 200 frames of 100 ms):

  * the trunk pitch and the image motion both oscillate with a period of 3.45 frames = half a gait cycle -> one gait cycle =
    0.69 s of recording = the 0.65 s wall-clock period: the GIF plays in real time;
  * after the goal the stride decays from full to rest over >= 3.6 s of recording (15 % .. 85 % span).  `brakes = 1 - (t - t_end)`
    reaches zero after at most 1.0 SIMULATED second (walk_env.py:237-244), so one simulated second took >= 3.6 wall seconds:
    the GUI playback ran at gait_clock_scale >= 3.6 (the ~265 pybullet API calls per control step through the GUI's
    shared-memory server do not fit a 5 ms budget; the training runs stored in the shipped checkpoint ran at ~16);
  * the trunk's pitch ripple in the recording is 0.42 deg rms (0.26 deg of that is measurement noise, see the backwards walk).
    The restated simulator at scale 4: 0.36 deg; at scale 1.5: 1.7 deg; at scale 1: falls.

So the reference is never observed at scale 1; at the clock it IS observed at, the restatement walks level to the goal like the
recording.  The bench keeps SURVEY's scale-1 workload as the headline (and now reports resets_per_step) and adds the demo-clock
workload next to it.
"""
import json
import os

import numpy as np

from oracle.oracle import OracleSim

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "readme_gif_series.json")))
DEMO_CLOCK = 4.0          # wall seconds per simulated second of the README walk_ik recording (>= 3.6 measured below)


def _smooth(x, n):
    return np.convolve(x, np.ones(n) / n, "same")


def _peak_period(x):
    d = x - _smooth(x, 13)
    d = d[20:-20]
    spec = np.abs(np.fft.rfft(d * np.hanning(len(d))))
    k = int(np.argmax(spec[3:])) + 3
    return len(d) / k, float(d.std())


def test_the_recording_plays_in_real_time():
    w = G["walk_ik"]
    assert w["frame_ms"] == [100] and w["frames"] == 200
    p_slope, _ = _peak_period(np.array(w["slope_deg"])[:140])
    p_motion, _ = _peak_period(np.array(w["motion"])[:140])
    # both series beat twice per gait cycle (one dip per diagonal pair): cycle = 2 x 3.45 frames x 0.1 s = 0.69 s ~ T = 0.65 s
    for p in (p_slope, p_motion):
        assert 0.55 < 2 * p * 0.1 < 0.80, p


def test_the_playback_ran_several_times_slower_than_real_time():
    m = _smooth(np.array(G["walk_ik"]["motion"]), 7)
    full, rest = np.median(m[60:130]), np.median(m[-12:])
    hi, lo = rest + 0.85 * (full - rest), rest + 0.15 * (full - rest)
    idx = np.arange(len(m))
    start = idx[(m >= hi) & (idx < 190)].max()            # last frame at full stride
    end = idx[(m <= lo) & (idx > start)].min()            # first frame at rest
    wall_seconds = (end - start) * 0.1
    assert wall_seconds >= 3.0, wall_seconds              # measured 3.6 s for the 15..85 % span alone
    # the whole ramp lasts at most 1.0 simulated second (brakes = 1 - (t - t_end) >= 0), the 15..85 % span 0.7 of it
    assert wall_seconds / 0.7 >= DEMO_CLOCK


def test_pitch_ripple_of_the_recorded_trot():
    _, rms = _peak_period(np.array(G["walk_ik"]["slope_deg"])[:140])
    _, noise = _peak_period(np.array(G["walk_back_ik"]["slope_deg"])[:140])
    assert 0.25 < rms < 0.6 and noise < rms                # 0.42 deg; the trunk stays level within a degree


def _walk(scale, n_envs=16, steps=2500, seed=0):
    s = OracleSim(n_envs, "walk", "ik", target_position=2.0, backwards=False, gait_clock_scale=scale)
    s.reset()
    rng = np.random.default_rng(seed)
    done_any = np.zeros(n_envs, bool)
    pitch = []
    for _ in range(steps):
        a = rng.uniform(-0.4, 0.4, (n_envs, 2)).astype(np.float32)
        a[:n_envs // 2] = 0.0                                   # half the batch: zero actions, half: random actions
        o, r, d = s.step(a, nthreads=4)
        done_any |= d
        pitch.append(o[:, 1].copy())
    x = np.array([s.state(i)["pos"][0] for i in range(n_envs)])
    goal = np.array([s.env(i).goal_reached for i in range(n_envs)])
    return x, goal, done_any, np.degrees(np.array(pitch))


def test_walk_ik_reaches_its_target_at_the_clock_the_reference_is_observed_at():
    """walk_env.py:252-290 + rex_gym_env.py:490-495: 16 envs, zero and random actions, forward to target 2.0 m."""
    x, goal, done_any, pitch = _walk(DEMO_CLOCK)
    assert not done_any.any() and goal.all()
    assert (np.abs(x) >= 1.85).all() and (np.abs(x) < 2.15).all()      # goal latch at |x| >= target - 0.15, then the brake
    ripple = pitch[300:800].std(0)
    assert ripple.max() < 0.6                                          # deg; recorded 0.42 incl. 0.26 of noise


def test_trot_stability_margin_of_the_restated_model():
    """The same walk for the clocks between the nominal one and the demo: everything from 1.5 up reaches the goal."""
    for scale in (1.5, 2.0, 3.0, 5.0):
        x, goal, done_any, _ = _walk(scale, n_envs=8)
        assert not done_any.any() and goal.all() and (np.abs(x) >= 1.85).all(), scale
    # scale 1 (SURVEY's bench workload): tips over after ~240 control steps in this model -- recorded here so a change of the
    # physics that moves the margin shows up
    x, goal, done_any, _ = _walk(1.0, n_envs=8, steps=400)
    assert done_any.all()
