from .batched_env import BatchedGoEnv
from .go_env import GoEnv, GoExtraHardEnv, RewardMethod

__all__ = ["BatchedGoEnv", "GoEnv", "GoExtraHardEnv", "RewardMethod"]
