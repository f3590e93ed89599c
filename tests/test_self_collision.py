"""The reference loads the URDF with URDF_USE_SELF_COLLISION (rex_gym/model/rex.py:275-281, rex_gym/envs/rex_gym_env.py:62): every
pair of links except parent-child collides.  The restated contact model carries no link-link rows; this test is why that is
the same physics on the path: along the trajectories of the tasks (random actions, falls and restarts included) no non-adjacent
pair of collision shapes comes closer than several millimetres -- eight times the 0.81 mm distance at which Bullet would open a
contact manifold.  The only pairs in permanent 'contact' are the chassis boxes and the shoulder boxes next to them (sibling
links, 0 and 1 mm apart in every pose): their faces are perpendicular to the shoulder axis, so they slide in their own plane and
never press on each other.  (tools/experiments/self_collision_survey.py holds the survey over static poses and all eight
task / signal pairs: closest approach 6.6 mm, turn-ol and standup.)"""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("self_collision_survey", os.path.join(ROOT, "tools", "experiments", "self_collision_survey.py"))
S = importlib.util.module_from_spec(spec)
spec.loader.exec_module(S)


def test_chassis_and_shoulder_boxes_touch_without_pressing():
    import numpy as np
    R, p = S.fk(S.pose(0.3, -1.0, 1.5))
    W = {n: (R[bi], pts @ R[bi].T + p[bi]) for n, bi, pts in S.shapes}
    for a, b in S.CONST:
        sep = S.separation(W[a][1], W[b][1], W[a][0], W[b][0])
        assert -1e-9 < sep < 1.5e-3                              # a 0 / 1 mm gap whatever the shoulder angle
        # the facing faces are x = const planes and the shoulder turns about x: the gap does not depend on the joint angle
    R2, p2 = S.fk(S.pose(-0.3, -1.0, 1.5))
    W2 = {n: (R2[bi], pts @ R2[bi].T + p2[bi]) for n, bi, pts in S.shapes}
    for a, b in S.CONST:
        assert abs(S.separation(W[a][1], W[b][1], W[a][0], W[b][0]) - S.separation(W2[a][1], W2[b][1], W2[a][0], W2[b][0])) < 1e-9


@pytest.mark.parametrize("task,sig,kw", [("walk", "ik", dict(target_position=2.0, backwards=False)), ("standup", "ol", {}), ("turn", "ol", {})])
def test_no_link_pair_comes_near_on_the_task_trajectories(task, sig, kw):
    sep, where = S.closest_pair_on_rollout(task, sig, kw, steps=120, n=4, every=6)
    assert sep > 4e-3, (sep, where)
