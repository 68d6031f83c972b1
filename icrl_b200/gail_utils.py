"""GailDiscriminator as a *cost source* for `cpg --load_gail` (icrl/cpg.py:54-75): load a discriminator saved by the
reference (icrl/gail_utils.py:316-393) and evaluate `reward_function` on the device.

Only inference is built: the discriminator is the same ReLU MLP + sigmoid as the constraint net, and its nominal-data path
(icrl/gail_utils.py:233-250) selects the input dimensions and does nothing else -- the reference's observation
normalisation / clipping and action clipping are commented out there -- so it runs on K1 (`icrl_cn_forward`, prediction
output).  Training the discriminator (the GAIL baseline, icrl/gail.py) is outside the ICRL hot path."""
from typing import Optional, Tuple

import numpy as np
import torch as th

from .constraint_net import ConstraintNet


class GailDiscriminator:
    def __init__(self, obs_dim: int, acs_dim: int, hidden_sizes: Tuple[int, ...], batch_size=None, lr_schedule=None,
                 expert_obs=None, expert_acs=None, is_discrete: bool = False, obs_select_dim=None, acs_select_dim=None,
                 optimizer_class=None, optimizer_kwargs=None, clip_obs: Optional[float] = 10., initial_obs_mean=None,
                 initial_obs_var=None, action_low=None, action_high=None, num_spurious_features=None,
                 freeze_weights: Optional[bool] = False, eps: float = 1e-5, device: str = "cpu"):
        if num_spurious_features is not None:
            raise NotImplementedError("spurious features are a GAIL-baseline experiment, not part of the cpg cost path")
        if optimizer_class is not None:
            raise NotImplementedError("training the discriminator (icrl/gail.py) is outside the ICRL hot path; "
                                      "GailDiscriminator.load gives an inference-only object")
        self.obs_dim, self.acs_dim, self.hidden_sizes, self.is_discrete = obs_dim, acs_dim, hidden_sizes, is_discrete
        self.obs_select_dim, self.acs_select_dim = obs_select_dim, acs_select_dim
        self.expert_obs, self.expert_acs = expert_obs, expert_acs
        # kept as attributes like the reference does, but NOT applied to nominal data (gail_utils.py:242-245)
        self.clip_obs, self.current_obs_mean, self.current_obs_var = clip_obs, initial_obs_mean, initial_obs_var
        self.action_low, self.action_high = action_low, action_high
        self.eps, self.device = eps, device
        self._net = ConstraintNet(obs_dim, acs_dim, hidden_sizes, None, None, None, None, is_discrete,
                                  obs_select_dim=obs_select_dim, acs_select_dim=acs_select_dim, optimizer_class=None,
                                  clip_obs=None, initial_obs_mean=None, initial_obs_var=None, action_low=None,
                                  action_high=None, eps=eps, device=device)
        self.select_dim, self.input_dims = self._net.select_dim, self._net.input_dims

    @property
    def network(self):
        return self._net.network

    def reward_function(self, obs: np.ndarray, acs: np.ndarray, apply_log: bool = True) -> np.ndarray:
        """gail_utils.py:146-156: D(s, a), or log(D + eps); `cpg` uses apply_log=False as the cost."""
        assert obs.shape[-1] == self.obs_dim, ""
        if not self.is_discrete:
            assert acs.shape[-1] == self.acs_dim, ""
        pred = self._net._forward_host(obs, acs, out_kind=1)
        if pred.ndim == 1:                      # the reference reshapes 2-D input to (n, 1) and squeezes it again
            pred = pred.reshape(-1, 1)
        if apply_log:
            return np.squeeze(np.log(pred + self.eps))
        return np.squeeze(pred)

    @classmethod
    def load(cls, load_path: str, obs_dim=None, acs_dim=None, is_discrete=None, expert_obs=None, expert_acs=None,
             obs_select_dim=None, acs_select_dim=None, clip_obs=None, obs_mean=None, obs_var=None, action_low=None,
             action_high=None, device: str = "auto"):
        state_dict = th.load(load_path, weights_only=False)
        pick = lambda v, k: state_dict[k] if v is None else v
        net = cls(pick(obs_dim, "obs_dim"), pick(acs_dim, "acs_dim"), state_dict["hidden_sizes"], None, None, expert_obs,
                  expert_acs, pick(is_discrete, "is_discrete"), pick(obs_select_dim, "obs_select_dim"),
                  pick(acs_select_dim, "acs_select_dim"), None, None, pick(clip_obs, "clip_obs"),
                  pick(obs_mean, "obs_mean"), pick(obs_var, "obs_var"), pick(action_low, "action_low"),
                  pick(action_high, "action_high"), device=state_dict["device"] if device is None else device)
        net._net.load_network_state_dict(state_dict["network"])
        return net
