"""Pack the URDF-derived tables (model/rex_<mark>.json, produced by tools/compile_urdf.py from
rex_gym/util/pybullet_data/assets/urdf/rex.urdf) into the flat float32 table the kernels TMA-load
(layout: include/rexsim.h REXSIM_MT_*)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MT_BASE, MT_LEG, MT_TOE, MT_BOX, MT_BASEBOX, MT_FLOATS = 0, 16, 208, 592, 880, 952
MAX_TOE_PTS = 32
TOE_MARGIN = 0.001   # Bullet's URDF importer puts a 1 mm collision margin on convex hulls


def _sym6(I):
    I = np.asarray(I, dtype=np.float64)
    return [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]]


def load_model_json(mark="base"):
    with open(os.path.join(HERE, "model", f"rex_{mark}.json")) as f:
        return json.load(f)


def pack_model_tables(mark="base"):
    """Returns (float32[MT_FLOATS], toe_npts).  Raises ValueError when the model does not have the leg
    structure the kernels assume (x-axis shoulder, y-axis leg and foot, unrotated joint frames)."""
    j = load_model_json(mark)
    bodies = j["bodies"]
    if len(bodies) < 13:
        raise ValueError("model must have a base and four 3-joint legs")
    t = np.zeros(MT_FLOATS, dtype=np.float64)
    b0 = bodies[0]
    t[MT_BASE + 0] = b0["mass"]
    t[MT_BASE + 1:MT_BASE + 4] = b0["com"]
    t[MT_BASE + 4:MT_BASE + 10] = _sym6(b0["inertia"])
    t[MT_BASE + 10] = j["root_mass"]
    t[MT_BASE + 11:MT_BASE + 14] = j["root_inertia"]
    boxes = [s for s in b0["shapes"] if s["kind"] == "box"]
    if len(boxes) != 3:
        raise ValueError("expected base + 2 chassis collision boxes")
    for k, s in enumerate(boxes):
        t[MT_BASEBOX + 24 * k:MT_BASEBOX + 24 * (k + 1)] = np.asarray(s["points"]).reshape(-1)
    axes = ([1, 0, 0], [0, 1, 0], [0, 1, 0])
    npts = None
    for leg in range(4):
        for bi in range(3):
            b = bodies[1 + 3 * leg + bi]
            if b["parent"] != (0 if bi == 0 else 3 * leg + bi):
                raise ValueError("unexpected kinematic tree")
            if not np.allclose(b["axis"], axes[bi]) or np.any(np.abs(b["joint_rpy"]) > 0):
                raise ValueError("kernels assume x/y/y joint axes and unrotated joint frames")
            o = MT_LEG + 48 * leg + 16 * bi
            t[o:o + 3] = b["joint_xyz"]
            t[o + 3] = b["mass"]
            t[o + 4:o + 7] = b["com"]
            t[o + 7] = b["lower"]
            t[o + 8:o + 14] = _sym6(b["inertia"])
            t[o + 14] = b["upper"]
            box = [s for s in b["shapes"] if s["kind"] == "box"]
            if len(box) != 1:
                raise ValueError("expected one collision box per leg body")
            ob = MT_BOX + 72 * leg + 24 * bi
            t[ob:ob + 24] = np.asarray(box[0]["points"]).reshape(-1)
            hull = [s for s in b["shapes"] if s["kind"] == "hull"]
            if bi == 2:
                if len(hull) != 1:
                    raise ValueError("expected one toe hull on the foot body")
                pts = np.asarray(hull[0]["points"])
                if npts is None:
                    npts = len(pts)
                if len(pts) != npts or npts > MAX_TOE_PTS:
                    raise ValueError("toe hulls must have the same number (<=32) of sample points")
                ot = MT_TOE + 3 * MAX_TOE_PTS * leg
                t[ot:ot + 3 * npts] = pts.reshape(-1)
    return t.astype(np.float32), int(npts)
