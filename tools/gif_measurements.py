#!/usr/bin/env python3
"""Measure the reference's own README animations (images/*.gif: screen recordings of the PyBullet GUI playing the shipped
policies, README.md:196-269) and commit the per-frame series as a small fixture, tests/golden/readme_gif_series.json.

Why: `GaitPlanner.loop` reads the WALL clock (rex_gym/model/gait_planner.py:108-110) while `_evaluate_gait_stage_coeff` /
`_evaluate_brakes_stage_coeff` (walk_env.py:228-244) run on the SIMULATION clock, so one recording contains both clocks:
  * the gait period in frames  -> the recording is real time (0.65 s gait = 6.5 frames of 100 ms),
  * the length of the braking ramp after the goal (`1 - (t - end_time)` for 0.8 + a1 <= 1.2 SIMULATED seconds) in frames
    -> how many wall seconds one simulated second took, i.e. the `gait_clock_scale` the demo ran at,
  * the pitch of the trunk (slope of its upper edge) -> the body ripple of a real PyBullet trot at that clock.
/root/reference exists only in the build container, hence the fixture (plus this script).  PIL is used to decode the GIFs.

Series per animation (one value per frame):
  motion   mean |frame - previous frame| over a window around the robot (the camera follows the base): gait activity
  slope    least-squares slope [deg] of the topmost yellow (trunk) pixel per column, 45 columns around the trunk centre
"""
import json
import os
import sys

import numpy as np
from PIL import Image

REF = "/root/reference/images"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "readme_gif_series.json")
NAMES = ["walk_ik", "walk_back_ik", "gallop_ik", "standup_ol"]


def series(name):
    im = Image.open(os.path.join(REF, name + ".gif"))
    prev, motion, slope, dur = None, [], [], []
    for i in range(im.n_frames):
        im.seek(i)
        dur.append(int(im.info.get("duration", 0)))
        rgb = np.asarray(im.convert("RGB")).astype(int)
        g = np.asarray(im.convert("L")).astype(float)[150:245, 250:350]
        motion.append(0.0 if prev is None else float(np.abs(g - prev).mean()))
        prev = g
        m = (rgb[..., 0] > 170) & (rgb[..., 1] > 150) & (rgb[..., 2] < 90)          # the yellow trunk / upper legs
        m[:, :60] = False; m[:, 340:] = False; m[:100] = False; m[260:] = False       # GUI panels, sky, floor reflections
        ys, xs = np.nonzero(m)
        if len(xs) < 50:
            slope.append(float("nan")); continue
        cx = int(np.median(xs))
        cols, tops = [], []
        for x in range(cx - 22, cx + 23):
            col = np.nonzero(m[:, x])[0]
            if len(col):
                cols.append(x); tops.append(col.min())
        cols, tops = np.array(cols), np.array(tops)
        p = np.polyfit(cols, tops, 1)
        ok = np.abs(tops - np.polyval(p, cols)) < 2
        p = np.polyfit(cols[ok], tops[ok], 1)
        slope.append(float(np.degrees(np.arctan(p[0]))))
    return {"frames": im.n_frames, "frame_ms": sorted(set(dur)), "size": list(im.size),
            "motion": [round(v, 3) for v in motion], "slope_deg": [round(v, 3) for v in slope]}


def main():
    out = {"source": "nicrusso7/rex-gym images/<name>.gif (README.md:196-269)", "script": "tools/gif_measurements.py"}
    for n in NAMES:
        out[n] = series(n)
        print(n, out[n]["frames"], "frames of", out[n]["frame_ms"], "ms")
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
