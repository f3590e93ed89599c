"""Same module path as rex_gym/envs/gym/walk_env.py."""
from ..single_env import RexWalkEnv  # noqa: F401
