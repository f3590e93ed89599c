"""Heightfield bank for terrain_type='random' (rex_gym/model/terrain.py:26,32-53,84-106).

The reference regenerates one private 256x256 field per env per reset from Python's global `random`
stream seeded with 10.  65 536 private fields do not fit, so the batched simulator keeps a read-only
bank: field k is the k-th 256x256 draw of that very stream (field 0 == the reference's first terrain),
and env e uses field (e + reset_count) mod nfields."""
import random

import numpy as np

ROWS = COLUMNS = 256
CELL = 0.05                 # meshScale (.05, .05, 1)
PERTURBATION = 0.05


def make_random_fields(nfields, seed=10):
    rnd = random.Random(seed)                       # terrain.py:26
    out = np.zeros((nfields, ROWS * COLUMNS), dtype=np.float32)
    for k in range(nfields):
        f = out[k]
        for j in range(COLUMNS // 2):               # terrain.py:36-44: 2x2 blocks share a height
            for i in range(ROWS // 2):
                h = rnd.uniform(0, PERTURBATION)
                f[2 * i + 2 * j * ROWS] = h
                f[2 * i + 1 + 2 * j * ROWS] = h
                f[2 * i + (2 * j + 1) * ROWS] = h
                f[2 * i + 1 + (2 * j + 1) * ROWS] = h
    return out.reshape(nfields, COLUMNS, ROWS)
