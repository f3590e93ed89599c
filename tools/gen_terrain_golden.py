#!/usr/bin/env python3
"""Golden for SURVEY row a17: the reference's own `Terrain.generate_terrain` / `update_terrain` (rex_gym/model/terrain.py:32-53,
84-106), imported from /root/reference and run UNMODIFIED with a recording pybullet stub -- what is kept is exactly what the
reference hands to `createCollisionShape(GEOM_HEIGHTFIELD, ...)`: the 65 536 heights of the first terrain (construction) and of
the second (first `update_terrain`, i.e. the next reset), mesh scale, rows / columns.

Output: tests/golden/terrain_golden.json (sha256 of the float32 heights + strided samples; the full arrays are 2 x 256 KB of noise).
"""
import hashlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_golden import install_stubs  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "terrain_golden.json")


def main():
    pb = install_stubs()
    calls = []

    def create_collision_shape(**kw):
        calls.append(kw)
        return len(calls)
    pb.GEOM_HEIGHTFIELD, pb.GEOM_CONCAVE_INTERNAL_EDGE, pb.COV_ENABLE_RENDERING = 9, 2, 7
    pb.createCollisionShape = create_collision_shape
    client = types.SimpleNamespace(GEOM_HEIGHTFIELD=9, COV_ENABLE_RENDERING=7, createCollisionShape=create_collision_shape,
                                   setAdditionalSearchPath=lambda *a: None, configureDebugVisualizer=lambda *a: None,
                                   createMultiBody=lambda *a, **k: 5, resetBasePositionAndOrientation=lambda *a: None,
                                   changeVisualShape=lambda *a, **k: None)
    from rex_gym.model import terrain as tmod
    t = tmod.Terrain("random", "random")
    t.generate_terrain(types.SimpleNamespace(pybullet_client=client))
    t.update_terrain()
    t.update_terrain()
    assert len(calls) == 3
    fields = [np.asarray(c["heightfieldData"], np.float64) for c in calls]
    out = {"source": "rex_gym/model/terrain.py:26,32-53,84-106 run unmodified; pybullet = recorder",
           "rows": calls[0]["numHeightfieldRows"], "columns": calls[0]["numHeightfieldColumns"], "mesh_scale": list(calls[0]["meshScale"]),
           "update_mesh_scale": list(calls[1]["meshScale"]), "update_replaces_shape": calls[1]["replaceHeightfieldIndex"],
           "init_position_random": tmod.ROBOT_INIT_POSITION["random"],
           "fields": [{"sha256_float32": hashlib.sha256(f.astype(np.float32).tobytes()).hexdigest(), "min": float(f.min()), "max": float(f.max()),
                       "every_257th": f[::257].tolist()} for f in fields]}
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes", [x["sha256_float32"][:12] for x in out["fields"]])


if __name__ == "__main__":
    main()
