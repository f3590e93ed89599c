"""Same module path as rex_gym/envs/gym/turn_env.py."""
from ..single_env import RexTurnEnv  # noqa: F401
