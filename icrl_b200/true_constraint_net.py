"""Ground-truth cost functions per environment id (mirror of icrl/true_constraint_net.py:11-60, 107-115): used by the
drivers for the `true/cost` metric and by `cpg` when no constraint net is given.  Host numpy, not on the hot path.
The bridge environments' cost needs the reference's `custom_envs` package and is resolved lazily."""
from functools import partial

import numpy as np


def wall_behind(pos, obs, acs):
    return (obs[..., 0] <= pos)


def wall_infront(pos, obs, acs):
    return (obs[..., 0] >= pos)


def wall_behind_and_infront(pos_back, pos_front, obs, acs):
    return (obs[..., 0] <= pos_back).astype(np.float32) + (obs[..., 0] >= pos_front).astype(np.float32)


def null_cost(x, *args):
    return np.zeros(x.shape[:1])


def torque_constraint(threshold, obs, acs):
    return np.any(np.abs(acs) > threshold, axis=-1)


def lap_grid_world(obs, acs):
    return np.array([1 if ac == 1 else 0 for ac in acs])


def get_true_cost_function(env_id):
    if env_id in ("HCWithPosTest-v0", "WalkerWithPosTest-v0", "SwimmerWithPosTest-v0", "AntWallTest-v0",
                  "AntWallBrokenTest-v0", "PointCircleTestBack-v0", "SynthHCWithPosTest-v0"):
        return partial(wall_behind, -3)
    if env_id in ("PointNullRewardTest-v0", "PointCircleTest-v0", "AntCircleTest-v0"):
        return partial(wall_behind_and_infront, -3, +3)
    if env_id in ("CLGW-v0", "SynthCLGW-v0"):
        return lap_grid_world
    if env_id in ("AntTest-v0", "HalfCheetahTest-v0", "Walker2dTest-v0", "SwimmerTest-v0"):
        return partial(torque_constraint, 0.5)
    if env_id in ("CDD2B-v0", "CC2B-v0", "CDD3B-v0"):
        raise NotImplementedError(f"the bridge cost of {env_id} lives in the reference's custom_envs package")
    print("Cost function for %s is not implemented yet. Returning null cost function" % env_id)
    return null_cost
