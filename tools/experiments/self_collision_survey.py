#!/usr/bin/env python3
"""Which link pairs of Rex can touch each other?  (reference: URDF_USE_SELF_COLLISION, rex_gym/model/rex.py:275-281,
rex_gym/envs/rex_gym_env.py:62: every pair of links except parent-child collides.)

Separation of two convex shapes = max over directions n of [min_A n.x - max_B n.x]; evaluated over a dense direction set
(face normals, edge cross products, 4000 sphere points), so a NEGATIVE value proves nothing smaller than the true penetration and a
positive value is a lower bound of the distance.  Poses: the task init poses, the folded rest pose, and a grid over the joint
ranges the gaits visit."""
import itertools, json, math, os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
j = json.load(open(os.path.join(ROOT, "rex_gym_b200", "model", "rex_base.json")))
bodies = j["bodies"]

def rot(axis, q):
    a = np.asarray(axis, float); c, s = math.cos(q), math.sin(q); t = 1 - c
    x, y, z = a
    return np.array([[t*x*x+c, t*x*y-s*z, t*x*z+s*y], [t*x*y+s*z, t*y*y+c, t*y*z-s*x], [t*x*z-s*y, t*y*z+s*x, t*z*z+c]])

def fk(q):
    R = [np.eye(3)]; p = [np.zeros(3)]
    for i, b in enumerate(bodies[1:], 1):
        pa = b["parent"]
        R.append(R[pa] @ rot(b["axis"], q[i - 1])); p.append(p[pa] + R[pa] @ np.asarray(b["joint_xyz"]))
    return R, p

shapes = []          # (link name, body index, points)
for i, b in enumerate(bodies):
    for sh in b["shapes"]:
        shapes.append((sh["link"], i, np.asarray(sh["points"], float)))
names = [s[0] for s in shapes]
def parent_child(a, b):
    fam = [("base_link", "chassis_front_link"), ("base_link", "chassis_rear_link")]
    for leg in ("front_left", "front_right", "rear_left", "rear_right"):
        fam += [("base_link", leg + "_shoulder_link"), (leg + "_shoulder_link", leg + "_leg_link"),
                (leg + "_leg_link", leg + "_foot_link"), (leg + "_foot_link", leg + "_toe_link")]
    return (a, b) in fam or (b, a) in fam
rng = np.random.default_rng(0)
D = rng.normal(size=(4000, 3)); D /= np.linalg.norm(D, axis=1, keepdims=True)
D = np.vstack([D, np.eye(3), -np.eye(3)])

def separation(PA, PB, RA, RB):
    dirs = [D] + [np.vstack([R.T, -R.T]) for R in (RA, RB)]
    ed = [np.cross(RA[:, a], RB[:, b]) for a in range(3) for b in range(3)]
    ed = np.array([e / np.linalg.norm(e) for e in ed if np.linalg.norm(e) > 1e-9])
    if len(ed): dirs.append(np.vstack([ed, -ed]))
    N = np.vstack(dirs)
    return float((np.min(PA @ N.T, axis=0) - np.max(PB @ N.T, axis=0)).max())

CONST = {("chassis_rear_link", "rear_left_shoulder_link"), ("chassis_rear_link", "rear_right_shoulder_link"),
         ("chassis_front_link", "front_left_shoulder_link"), ("chassis_front_link", "front_right_shoulder_link")}


def closest_pair_on_rollout(task, sig, kw, steps=300, n=6, every=4, seed=1):
    """Smallest separation of any non-adjacent link pair (the four chassis/shoulder face pairs aside: constant 0 / 1 mm gap, faces
    that slide in their own plane) along an oracle rollout on random actions, episodes restarting when they end."""
    sys.path.insert(0, ROOT)
    from oracle.oracle import OracleSim
    global D
    Dsave, D = D, D[::3]
    o = OracleSim(n, task, sig, normalize=True, **kw)
    o.reset()
    rngs = np.random.default_rng(seed)
    best = (9.0, None)
    for k in range(steps):
        obs, r, d = o.step(rngs.uniform(-1, 1, (n, o.A)).astype(np.float32), 6)
        if k % every == 0:
            for i in range(n):
                R, p = fk(o.state(i)["q"])
                W = [(R[bi], pts @ R[bi].T + p[bi]) for _, bi, pts in shapes]
                for a, b in itertools.combinations(range(len(shapes)), 2):
                    if parent_child(names[a], names[b]) or (names[a], names[b]) in CONST: continue
                    if np.linalg.norm(W[a][1].mean(0) - W[b][1].mean(0)) > 0.16: continue      # far apart: skip the exact test
                    sep = separation(W[a][1], W[b][1], W[a][0], W[b][0])
                    if sep < best[0]: best = (sep, (names[a], names[b], k))
        idx = np.nonzero(d)[0]
        if len(idx): o.reset(idx)
    D = Dsave
    return best


def survey(label, qs):
    worst = {}
    for q in qs:
        R, p = fk(q)
        W = [(R[bi], pts @ R[bi].T + p[bi]) for _, bi, pts in shapes]
        for a, b in itertools.combinations(range(len(shapes)), 2):
            if parent_child(names[a], names[b]): continue
            s = separation(W[a][1], W[b][1], W[a][0], W[b][0])
            k = (names[a], names[b])
            if k not in worst or s < worst[k]: worst[k] = s
    close = sorted((v, k) for k, v in worst.items() if v < 0.004)
    print("== %s: %d poses, %d pairs checked, %d within 4 mm" % (label, len(qs), len(worst), len(close)))
    for v, k in close[:14]: print("   %+.4f m  %s  x  %s" % (v, k[0], k[1]))

def pose(sh, lg, ft):
    return np.array([sh, lg, ft, -sh, lg, ft, sh, lg, ft, -sh, lg, ft])


TASKS = (("walk", "ik", dict(target_position=2.0, backwards=False)), ("walk", "ol", {}), ("gallop", "ik", {}), ("gallop", "ol", {}),
         ("turn", "ik", {}), ("turn", "ol", {}), ("standup", "ol", {}), ("poses", "ik", {}))

if __name__ == "__main__":
    survey("stand (ik init)", [pose(0, -0.88643435, 1.30197369)])
    survey("stand_ol (ol init)", [pose(0.15192765, -0.90412283, 1.48156545)])
    survey("rest_position at the foot limit", [pose(-0.4, -1.5, 2.59)])
    grid = [pose(s, l, f) for s in (-0.3, 0.0, 0.3) for l in np.linspace(-1.6, -0.4, 7) for f in np.linspace(0.7, 2.2, 7)]
    survey("gait range, legs in phase", grid)
    g2 = []
    for l1, f1, l2, f2 in itertools.product(np.linspace(-1.5, -0.5, 5), np.linspace(0.8, 2.0, 5), np.linspace(-1.5, -0.5, 5), np.linspace(0.8, 2.0, 5)):
        q = pose(0, l1, f1); q[6:9] = [0, l2, f2]; q[9:12] = [0, l2, f2]; g2.append(q)
    survey("gait range, front vs rear legs independent", g2[::7])

    if "--rollouts" in sys.argv:
        for task, sig, kw in TASKS:
            best = closest_pair_on_rollout(task, sig, kw)
            print("%-8s %-3s closest non-adjacent pair over 300 steps x 6 envs: %+.4f m  %s" % (task, sig, best[0], best[1]))
