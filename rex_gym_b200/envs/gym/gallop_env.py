"""Same module path as rex_gym/envs/gym/gallop_env.py."""
from ..single_env import RexReactiveEnv  # noqa: F401
