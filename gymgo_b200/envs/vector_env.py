"""GoVectorEnv - gymnasium.vector-style adapter over BatchedGoEnv.

Follows the gymnasium VectorEnv calling convention without importing gymnasium (not installed in this image):
    obs, info = venv.reset(seed=...)
    obs, reward, terminated, truncated, info = venv.step(actions)
with next-step autoreset: a sub-environment that terminated on step t is reset at the beginning of step t+1
(its action of step t+1 is then applied to the fresh board, as BatchedGoEnv.step(auto_reset=True) does).
All arrays are CUDA tensors; `info["action_mask"]` is the [B, N*N+1] legal-move mask of the new state."""
import torch

from .batched_env import BatchedGoEnv


class GoVectorEnv(object):
    def __init__(self, num_envs, size, komi=0, reward_method="real", device=None, obs_dtype=torch.float32, seed=0):
        self.num_envs, self.size = int(num_envs), int(size)
        self.env = BatchedGoEnv(num_envs, size, komi=komi, reward_method=reward_method, device=device,
                                obs_dtype=obs_dtype, strict=False, seed=seed)
        self.single_observation_shape = (6, size, size)
        self.single_action_n = size * size + 1
        self._truncated = torch.zeros(self.num_envs, dtype=torch.bool, device=self.env.rec.device)

    def _info(self):
        return {"action_mask": self.env.valid_moves(dtype=torch.uint8), "turn": self.env.turn()}

    def reset(self, seed=None, options=None):
        if seed is not None:
            self.env.seed, self.env.t = int(seed), 0
        obs = self.env.reset()
        return obs, self._info()

    def step(self, actions):
        obs, reward, done, info = self.env.step(actions, auto_reset=True)
        out = self._info()
        out["status"] = info["status"]            # non-zero where an illegal action was refused (board unchanged)
        return obs, reward, done.bool(), self._truncated, out

    def sample_actions(self):
        """uniformly random legal actions for every sub-environment (GoEnv.uniform_random_action)"""
        return self.env.uniform_random_action()

    def close(self):
        pass
