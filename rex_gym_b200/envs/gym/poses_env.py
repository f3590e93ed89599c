"""Same module path as rex_gym/envs/gym/poses_env.py."""
from ..single_env import RexPosesEnv  # noqa: F401
