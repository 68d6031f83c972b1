"""Sample containers with the reference's field order (stable_baselines3/common/type_aliases.py:28-38)."""
from typing import NamedTuple

import torch as th


class RolloutBufferWithCostSamples(NamedTuple):
    orig_observations: th.Tensor
    observations: th.Tensor
    actions: th.Tensor
    old_log_prob: th.Tensor
    old_reward_values: th.Tensor
    reward_advantages: th.Tensor
    reward_returns: th.Tensor
    old_cost_values: th.Tensor
    cost_advantages: th.Tensor
    cost_returns: th.Tensor
