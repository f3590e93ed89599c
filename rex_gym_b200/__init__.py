"""rex_gym_b200 -- B200-native batched replacement of rex-gym's per-step hot path
(RexGymEnv.step -> Rex.ApplyAction -> pybullet.stepSimulation -> Rex.GetObservation + gait controller)."""
from .envs.batched_env import (BatchedRexEnv, RexWalkBatchEnv, RexGallopBatchEnv, RexTurnBatchEnv,  # noqa: F401
                               RexStandupBatchEnv, RexPosesBatchEnv, make, ENV_IDS)

__all__ = ["BatchedRexEnv", "RexWalkBatchEnv", "RexGallopBatchEnv", "RexTurnBatchEnv", "RexStandupBatchEnv", "RexPosesBatchEnv", "make", "ENV_IDS"]
