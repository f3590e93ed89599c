"""Minimal stand-in for gym.spaces.Box (gym is not a dependency of the hot path; when gym is importable
the real class is used so `env.action_space.contains` etc. behave as the reference agent expects)."""
import numpy as np

try:  # pragma: no cover - gym is absent in the build image
    from gym.spaces import Box  # type: ignore
except Exception:
    class Box(object):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low = np.asarray(low, dtype=dtype)
            self.high = np.asarray(high, dtype=dtype)
            self.shape = self.low.shape
            self.dtype = np.dtype(dtype)

        def contains(self, x):
            x = np.asarray(x)
            # the reference's gallop Box has low > high (gallop_env.py:128-130): mirror gym, which then
            # rejects everything; BatchedRexEnv only checks shapes and finiteness for that reason
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __eq__(self, other):
            return isinstance(other, Box) and np.allclose(self.low, other.low) and np.allclose(self.high, other.high)

        def __repr__(self):
            return f"Box({self.low}, {self.high})"
