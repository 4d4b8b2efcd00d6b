from .batched_env import BatchedGoEnv
from .go_env import GoEnv, GoExtraHardEnv, RewardMethod
from .vector_env import GoVectorEnv

__all__ = ["BatchedGoEnv", "GoEnv", "GoExtraHardEnv", "GoVectorEnv", "RewardMethod"]
