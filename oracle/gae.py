"""TEST INFRASTRUCTURE (oracle) -- dual reward/cost GAE (K3).

CPU restatement of RolloutBufferWithCost._compute_returns_and_advantage /
compute_returns_and_advantage (/root/reference/stable_baselines3/common/buffers.py:493-552).
Pinned against the unmodified reference (tests/golden/gae_*.npz).

dtype trap reproduced on purpose: `last_dones` is a bool array, so
`1.0 - last_dones` is float64 and the running carry is float64 for the whole
reverse scan; `delta` is float32 for every step but the last; each advantage is
rounded to float32 when stored.  The order of every multiplication below is the
reference's (numpy rounds after each one), do not "simplify" it.
"""
import numpy as np


def gae_single(rewards, values, dones, gamma, gae_lambda, last_value, last_dones):
    """One reverse scan (buffers.py:526-541).

    rewards/values/dones: [T, E] float32; last_value [E] float32; last_dones [E] bool.
    Returns (returns, advantages), both [T, E] float32.
    """
    gamma, gae_lambda = float(gamma), float(gae_lambda)   # python floats are "weak" scalars in numpy promotion
    T = rewards.shape[0]
    adv = np.zeros(rewards.shape, dtype=np.float32)
    bootstrap = np.asarray(last_value, dtype=np.float32).flatten()

    # t = T-1 bootstraps from the value of the observation after the rollout; the
    # mask comes from a bool array, hence float64 from here on for the carry.
    alive = 1.0 - last_dones
    td = rewards[T - 1] + gamma * bootstrap * alive - values[T - 1]
    carry = td + gamma * gae_lambda * alive * 0
    adv[T - 1] = carry
    for t in range(T - 2, -1, -1):
        alive = 1.0 - dones[t + 1]                                   # float32
        td = rewards[t] + gamma * values[t + 1] * alive - values[t]  # float32, rounded per op
        carry = td + gamma * gae_lambda * alive * carry              # float64
        adv[t] = carry
    return adv + values, adv


def dual_gae(rewards, reward_values, costs, cost_values, dones, reward_last_value, cost_last_value, last_dones,
             reward_gamma, reward_gae_lambda, cost_gamma, cost_gae_lambda):
    """buffers.py:543-552: the same scan for (rewards, reward_values) and (costs, cost_values)."""
    rr, ra = gae_single(rewards, reward_values, dones, reward_gamma, reward_gae_lambda, reward_last_value, last_dones)
    cr, ca = gae_single(costs, cost_values, dones, cost_gamma, cost_gae_lambda, cost_last_value, last_dones)
    return {"reward_returns": rr, "reward_advantages": ra, "cost_returns": cr, "cost_advantages": ca}


def env_major(arr):
    """buffers.py:52-65 (`swap_and_flatten`): [T, E, ...] -> [E*T, ...], row = e*T + t."""
    if arr.ndim < 3:
        arr = arr[..., None]
    T, E = arr.shape[:2]
    return arr.swapaxes(0, 1).reshape(T * E, *arr.shape[2:])
