"""Same module path as rex_gym/envs/gym/standup_env.py."""
from ..single_env import RexStandupEnv  # noqa: F401
