from gymgo_b200.envs import GoEnv, GoExtraHardEnv, RewardMethod, BatchedGoEnv  # noqa: F401
