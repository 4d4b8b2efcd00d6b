from gymgo_b200.envs import GoEnv, GoExtraHardEnv, RewardMethod, BatchedGoEnv, GoVectorEnv  # noqa: F401
