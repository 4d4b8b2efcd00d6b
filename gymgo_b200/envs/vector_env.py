"""GoVectorEnv - gymnasium.vector-style adapter over BatchedGoEnv.

Follows the gymnasium VectorEnv calling convention without importing gymnasium (not installed in this image):
    obs, info = venv.reset(seed=...)
    obs, reward, terminated, truncated, info = venv.step(actions)

Autoreset (`autoreset_mode`):
  "next_step" (default, gymnasium's AutoresetMode.NEXT_STEP): a sub-environment that terminated on step t is reset by
      step t+1; the action passed for it on step t+1 is IGNORED, the step returns the reset observation, reward 0 and
      terminated False.
  "play_on_reset" (not a gymnasium mode): the reset happens at the start of step t+1 and the action of step t+1 is
      played on the fresh board - no step is spent on the reset; what the throughput benchmark's loop does.
Either way the whole step is ONE kernel launch (gg_step with GG_STEP_AUTO_RESET).

All arrays are CUDA tensors; `info["action_mask"]` is the [B, N*N+1] legal-move mask of the new state.
The observation / reward tensors returned by step() are the environment's static buffers and are overwritten by the
next step(); pass copy=True (or clone() them) before storing them in a replay buffer."""
import torch

from .batched_env import BatchedGoEnv


class GoVectorEnv(object):
    def __init__(self, num_envs, size, komi=0, reward_method="real", device=None, obs_dtype=torch.float32, seed=0,
                 autoreset_mode="next_step", copy=False):
        if autoreset_mode not in ("next_step", "play_on_reset"):
            raise ValueError("autoreset_mode must be 'next_step' or 'play_on_reset'")
        self.num_envs, self.size = int(num_envs), int(size)
        self.autoreset_mode, self.copy = autoreset_mode, bool(copy)
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
        return (obs.clone() if self.copy else obs), self._info()

    def step(self, actions):
        mode = "skip" if self.autoreset_mode == "next_step" else True
        obs, reward, done, info = self.env.step(actions, auto_reset=mode)
        out = self._info()
        out["status"] = info["status"]            # non-zero where an illegal action was refused (board unchanged)
        if self.copy:
            obs, reward, out["status"] = obs.clone(), reward.clone(), out["status"].clone()
        return obs, reward, done.bool(), self._truncated, out

    def sample_actions(self):
        """uniformly random legal actions for every sub-environment (GoEnv.uniform_random_action)"""
        return self.env.uniform_random_action()

    def close(self):
        pass
