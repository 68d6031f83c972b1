"""ConstraintNet -- host-side mirror of the reference class (icrl/constraint_net.py:14-402) whose arithmetic
runs in the sm_100a kernels of libicrl_b200.so (K1: cost_function, K2: train).

Same constructor signature, attributes (`select_dim`, `network`, `optimizer`, `current_obs_mean/var`, ...),
`cost_function`, `train`, `save`, `load` and checkpoint format as the reference, including its quirks
(positional-argument shift in `load`, per-step IS broadcast) -- see DESIGN.md.  Parameters and Adam moments
live in flat float32 device tensors; `network` / `optimizer` are thin views that import/export the
reference's state_dict formats.
"""
import ctypes as C
from collections import OrderedDict
from typing import Any, Callable, Dict, Optional, Tuple

import numpy as np
import torch as th
from torch import nn

from . import _lib
from .device import resolve_device


def create_mlp(input_dim, output_dim, net_arch, activation_fn=nn.ReLU):
    """Same module sequence (hence same state_dict keys and RNG consumption at init) as torch_layers.py:93-126."""
    mods, last = [], input_dim
    for width in net_arch:
        mods += [nn.Linear(last, width), activation_fn()]
        last = width
    if output_dim > 0:
        mods.append(nn.Linear(last, output_dim))
    return mods


class _FlatAdam:
    """The slice of th.optim.Adam's interface the reference touches (param_groups[..]['lr'], state_dict,
    load_state_dict), backed by flat device buffers that the CUDA Adam step updates in place."""

    def __init__(self, owner, lr, eps=1e-5, betas=(0.9, 0.999), **unused):
        self._owner = owner
        self.param_groups = [dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=0, amsgrad=False)]
        self.step_count = 0

    def state_dict(self):
        state = {}
        if self.step_count > 0:
            for i, (off, shape) in enumerate(self._owner._param_slices()):
                n = int(np.prod(shape))
                state[i] = dict(step=th.tensor(float(self.step_count)),
                                exp_avg=self._owner._adam_m[off:off + n].reshape(shape).cpu().clone(),
                                exp_avg_sq=self._owner._adam_v[off:off + n].reshape(shape).cpu().clone())
        group = dict(self.param_groups[0], params=list(range(len(self._owner._param_slices()))))
        return dict(state=state, param_groups=[group])

    def load_state_dict(self, sd):
        g = sd["param_groups"][0]
        for k in ("lr", "eps", "betas"):
            if k in g:
                self.param_groups[0][k] = tuple(g[k]) if k == "betas" else g[k]
        # torch >= 1.? numbers the state 0..n-1; older checkpoints (the reference's expert zips) key it by id(param)
        # and give the parameter order in param_groups[0]['params']
        keys = [k for k in g.get("params", []) if k in sd["state"]] or sorted(sd["state"].keys())
        for i, ((off, shape), k) in enumerate(zip(self._owner._param_slices(), keys)):
            n = int(np.prod(shape))
            st = sd["state"][k]
            self._owner._adam_m[off:off + n] = st["exp_avg"].reshape(-1).to(self._owner._adam_m)
            self._owner._adam_v[off:off + n] = st["exp_avg_sq"].reshape(-1).to(self._owner._adam_v)
            self.step_count = int(st["step"])


class ConstraintNet:
    def __init__(
            self,
            obs_dim: int,
            acs_dim: int,
            hidden_sizes: Tuple[int, ...],
            batch_size: int,
            lr_schedule: Callable[[float], float],
            expert_obs: np.ndarray,
            expert_acs: np.ndarray,
            is_discrete: bool,
            regularizer_coeff: float = 0.,
            obs_select_dim: Optional[Tuple[int, ...]] = None,
            acs_select_dim: Optional[Tuple[int, ...]] = None,
            optimizer_class=th.optim.Adam,
            optimizer_kwargs: Optional[Dict[str, Any]] = None,
            no_importance_sampling: bool = False,
            per_step_importance_sampling: bool = False,
            clip_obs: Optional[float] = 10.,
            initial_obs_mean: Optional[np.ndarray] = None,
            initial_obs_var: Optional[np.ndarray] = None,
            action_low: Optional[float] = None,
            action_high: Optional[float] = None,
            target_kl_old_new: float = -1,
            target_kl_new_old: float = -1,
            train_gail_lambda: Optional[bool] = False,
            eps: float = 1e-5,
            device: str = "cuda"
    ):
        self.obs_dim, self.acs_dim = obs_dim, acs_dim
        self.obs_select_dim, self.acs_select_dim = obs_select_dim, acs_select_dim
        self._define_input_dims()
        self.expert_obs, self.expert_acs = expert_obs, expert_acs
        self.hidden_sizes = hidden_sizes
        self.batch_size = batch_size
        self.is_discrete = is_discrete
        self.regularizer_coeff = regularizer_coeff
        self.importance_sampling = not no_importance_sampling
        self.per_step_importance_sampling = per_step_importance_sampling
        self.clip_obs = clip_obs
        self.device = device
        self._dev = resolve_device(device)
        self.eps = eps
        self.train_gail_lambda = train_gail_lambda
        if optimizer_kwargs is None:
            optimizer_kwargs = {}
            if optimizer_class == th.optim.Adam:
                optimizer_kwargs["eps"] = 1e-5          # constraint_net.py:66-70
        if optimizer_class not in (None, th.optim.Adam):
            raise NotImplementedError("icrl_b200 implements the reference's optimiser (Adam) only")
        self.optimizer_kwargs, self.optimizer_class = optimizer_kwargs, optimizer_class
        self.lr_schedule = lr_schedule
        self.current_obs_mean, self.current_obs_var = initial_obs_mean, initial_obs_var
        self.action_low, self.action_high = action_low, action_high
        self.target_kl_old_new, self.target_kl_new_old = target_kl_old_new, target_kl_new_old
        self.current_progress_remaining = 1.
        self._expert_dev = None
        self._build()

    # ---------------------------------------------------------------- structure
    def _define_input_dims(self) -> None:
        """constraint_net.py:87-99 (acs_select_dim indexes concat([obs, acs]) without an obs_dim offset)."""
        sel = []
        if self.obs_select_dim is None:
            sel += list(range(self.obs_dim))
        elif self.obs_select_dim[0] != -1:
            sel += list(self.obs_select_dim)
        if self.acs_select_dim is None:
            sel += list(range(self.acs_dim))
        elif self.acs_select_dim[0] != -1:
            sel += list(self.acs_select_dim)
        assert len(sel) > 0, ""
        self.select_dim = sel
        self.input_dims = len(sel)

    def _build(self) -> None:
        hidden = tuple(int(h) for h in self.hidden_sizes)
        if not 1 <= len(hidden) <= _lib.MAX_HIDDEN or max(hidden) > _lib.CN_MAX_WIDTH:
            raise NotImplementedError(f"constraint net {hidden}: the CUDA kernels support 1..{_lib.MAX_HIDDEN} hidden "
                                      f"layers of width <= {_lib.CN_MAX_WIDTH}")
        if self.input_dims > _lib.MAX_SELECT:
            raise NotImplementedError(f"more than {_lib.MAX_SELECT} selected inputs")
        # a CPU module only for default-init RNG parity and state_dict key names; never used for compute
        self._shape_net = nn.Sequential(*create_mlp(self.input_dims, 1, hidden), nn.Sigmoid())
        self._slices, off = [], 0
        for p in self._shape_net.parameters():
            self._slices.append((off, tuple(p.shape)))
            off += p.numel()
        self._n_params = off
        self._params = th.cat([p.detach().reshape(-1) for p in self._shape_net.parameters()]).to(self._dev)
        self._adam_m = th.zeros_like(self._params)
        self._adam_v = th.zeros_like(self._params)
        self._stats_dirty = True
        self._desc = None
        if self.optimizer_class is not None:
            self.optimizer = _FlatAdam(self, lr=self.lr_schedule(1), **self.optimizer_kwargs)
        else:
            self.optimizer = None

    def _param_slices(self):
        return self._slices

    @property
    def network(self) -> nn.Sequential:
        """CPU copy of the net in the reference's module structure (state_dict keys 0.weight, 0.bias, 2.weight...)."""
        flat = self._params.detach().cpu()
        with th.no_grad():
            for p, (off, shape) in zip(self._shape_net.parameters(), self._slices):
                p.copy_(flat[off:off + p.numel()].reshape(shape))
        return self._shape_net

    def load_network_state_dict(self, sd) -> None:
        self._shape_net.load_state_dict(sd)
        self._params.copy_(th.cat([p.detach().reshape(-1) for p in self._shape_net.parameters()]))

    def parameters_flat(self) -> th.Tensor:
        return self._params

    # ---------------------------------------------------------------- C-ABI descriptor
    def _make_desc(self) -> _lib.CnDesc:
        d = _lib.CnDesc()
        d.obs_dim, d.acs_dim, d.is_discrete = self.obs_dim, self.acs_dim, int(bool(self.is_discrete))
        d.n_select = self.input_dims
        for i, s in enumerate(self.select_dim):
            d.select[i] = int(s)
        d.n_hidden = len(self.hidden_sizes)
        for i, h in enumerate(self.hidden_sizes):
            d.hidden[i] = int(h)
        mean, var = self.current_obs_mean, self.current_obs_var
        d.has_norm = int(mean is not None and var is not None)
        if d.has_norm:
            mean64 = np.asarray(mean, dtype=np.float64).reshape(-1)
            rstd64 = 1.0 / np.sqrt(np.asarray(var, dtype=np.float64).reshape(-1) + self.eps)
            self._mean_dev = th.from_numpy(np.ascontiguousarray(mean64)).to(self._dev)
            self._rstd_dev = th.from_numpy(np.ascontiguousarray(rstd64)).to(self._dev)
            d.obs_mean, d.obs_rstd = self._mean_dev.data_ptr(), self._rstd_dev.data_ptr()
        if self.clip_obs is not None and np.ndim(self.clip_obs) != 0:
            raise NotImplementedError("array-valued clip_obs (only reachable through load()'s argument shift with a "
                                      "non-None obs_mean) is not supported")
        d.has_clip_obs = int(self.clip_obs is not None)
        d.clip_obs = float(self.clip_obs) if self.clip_obs is not None else 0.0
        d.has_clip_acs = int(self.action_high is not None and self.action_low is not None and not self.is_discrete)
        if d.has_clip_acs:
            low = np.broadcast_to(np.asarray(self.action_low, dtype=np.float32), (self.acs_dim,)).copy()
            high = np.broadcast_to(np.asarray(self.action_high, dtype=np.float32), (self.acs_dim,)).copy()
            self._low_dev = th.from_numpy(np.ascontiguousarray(low)).to(self._dev)
            self._high_dev = th.from_numpy(np.ascontiguousarray(high)).to(self._dev)
            d.acs_low, d.acs_high = self._low_dev.data_ptr(), self._high_dev.data_ptr()
        d.params = self._params.data_ptr()
        return d

    def _get_desc(self) -> _lib.CnDesc:
        if self._desc is None or self._stats_dirty:
            self._desc = self._make_desc()
            self._stats_dirty = False
        return self._desc

    def __setattr__(self, name, value):
        # any change to the normalisation / clipping attributes invalidates the cached descriptor
        if name in ("current_obs_mean", "current_obs_var", "clip_obs", "action_low", "action_high"):
            object.__setattr__(self, "_stats_dirty", True)
        object.__setattr__(self, name, value)

    # ---------------------------------------------------------------- host <-> device input marshalling
    def _host_inputs(self, obs: np.ndarray, acs: np.ndarray):
        obs = np.asarray(obs)
        if obs.dtype not in (np.float32, np.float64):
            obs = obs.astype(np.float64)
        lead = obs.shape[:-1]
        obs2 = np.ascontiguousarray(obs.reshape(-1, self.obs_dim))
        acs = np.asarray(acs)
        if self.is_discrete:
            acs2 = np.ascontiguousarray(acs.astype(np.int64).reshape(-1).astype(np.float32))
        else:
            acs2 = np.ascontiguousarray(acs.reshape(-1, self.acs_dim).astype(np.float32, copy=False))
        assert acs2.shape[0] == obs2.shape[0], "obs / acs row count mismatch"
        return obs2, acs2, lead

    # ---------------------------------------------------------------- K1
    def cost_function(self, obs: np.ndarray, acs: np.ndarray) -> np.ndarray:
        """constraint_net.py:121-130: cost = 1 - zeta(prepare_data(obs, acs)); host float32, shape obs.shape[:-1]."""
        assert obs.shape[-1] == self.obs_dim, ""
        if not self.is_discrete:
            assert acs.shape[-1] == self.acs_dim, ""
        return self._forward_host(obs, acs, out_kind=0)

    def _forward_host(self, obs, acs, out_kind):
        obs2, acs2, lead = self._host_inputs(obs, acs)
        out = np.empty(obs2.shape[0], dtype=np.float32)
        args = (C.byref(self._get_desc()), _lib.ptr(obs2), int(obs2.dtype == np.float64), _lib.ptr(acs2), obs2.shape[0],
                _lib.ptr(out), out_kind)
        # (this is the per-environment-step call of the cost wrapper: skip the device context manager when the net's
        #  device is already current -- it costs ~5 us of a ~30 us call)
        if th.cuda.current_device() == self._dev.index:
            _lib.check(_lib.lib().icrl_cn_forward_host(*args, _lib.current_stream()))
        else:
            with th.cuda.device(self._dev):
                _lib.check(_lib.lib().icrl_cn_forward_host(*args, _lib.current_stream()))
        return out.reshape(lead)

    def cost_function_device(self, obs: th.Tensor, acs: th.Tensor, out: Optional[th.Tensor] = None) -> th.Tensor:
        """K1 on device-resident buffers (whole-rollout relabel): obs [..., obs_dim] float32/float64 cuda tensor,
        acs [..., acs_dim] float32 (or action indices as float32 when discrete).  Asynchronous on the current stream."""
        assert obs.is_cuda and acs.is_cuda and obs.shape[-1] == self.obs_dim
        obs2 = obs.reshape(-1, self.obs_dim).contiguous()
        acs2 = (acs.reshape(-1) if self.is_discrete else acs.reshape(-1, self.acs_dim)).to(th.float32).contiguous()
        n = obs2.shape[0]
        if out is None:
            out = th.empty(n, dtype=th.float32, device=obs.device)
        with th.cuda.device(obs.device):
            _lib.check(_lib.lib().icrl_cn_forward(
                C.byref(self._get_desc()), _lib.ptr(obs2), int(obs2.dtype == th.float64), _lib.ptr(acs2), n,
                _lib.ptr(out), 0, _lib.current_stream()))
        return out.reshape(obs.shape[:-1])

    def call_forward(self, x: np.ndarray):
        raise NotImplementedError("call_forward on pre-selected inputs is only used by the reference's plotting code")

    # ---------------------------------------------------------------- K2
    def _update_learning_rate(self, current_progress_remaining) -> None:
        self.current_progress_remaining = current_progress_remaining
        for g in self.optimizer.param_groups:                # stable_baselines3/common/utils.py:62-71
            g["lr"] = self.lr_schedule(current_progress_remaining)

    def _expert_on_device(self, obs_dtype):
        """The expert batch is uploaded once and stays resident in HBM across ICRL iterations."""
        if self._expert_dev is None or self._expert_dev[0].dtype != obs_dtype:
            eo, ea, _ = self._host_inputs(self.expert_obs, self.expert_acs)
            eo_dev = th.from_numpy(eo).to(self._dev).to(obs_dtype)
            self._expert_dev = (eo_dev, th.from_numpy(ea).to(self._dev))
        return self._expert_dev

    def train(
            self,
            iterations: int,
            nominal_obs: np.ndarray,
            nominal_acs: np.ndarray,
            episode_lengths: np.ndarray,
            obs_mean: Optional[np.ndarray] = None,
            obs_var: Optional[np.ndarray] = None,
            current_progress_remaining: float = 1,
    ) -> Dict[str, Any]:
        """constraint_net.py:137-229 (full-batch mode).  One C-ABI call runs all `iterations` Adam steps on the
        device, with IS weights, both KLs and the early-stop test evaluated there; returns the backward/* metrics."""
        self._update_learning_rate(current_progress_remaining)
        self.current_obs_mean, self.current_obs_var = obs_mean, obs_var
        no, na, _ = self._host_inputs(nominal_obs, nominal_acs)
        lengths = np.asarray(episode_lengths, dtype=np.int64).reshape(-1)
        assert int(lengths.sum()) <= no.shape[0]
        offsets = np.zeros(len(lengths) + 1, dtype=np.int32)
        offsets[1:] = np.cumsum(lengths)
        exp_is_f64 = np.asarray(self.expert_obs).dtype != np.float32
        if exp_is_f64 and no.dtype != np.float64:
            no = no.astype(np.float64)
        eo_dev, ea_dev = self._expert_on_device(th.float64 if no.dtype == np.float64 else th.float32)
        assert int(lengths.sum()) == no.shape[0] or not self.importance_sampling, "episode_lengths must cover nominal_obs"
        no_dev, na_dev = th.from_numpy(no).to(self._dev), th.from_numpy(na).to(self._dev)
        off_dev = th.from_numpy(offsets).to(self._dev)
        comm = getattr(self, "comm", None)
        dp = comm is not None and comm.world > 1
        if dp:          # this rank's slice of the (replicated) expert batch; the nominal episodes are this rank's own
            n_exp = eo_dev.shape[0]
            lo, hi = n_exp * comm.rank // comm.world, n_exp * (comm.rank + 1) // comm.world
            eo_dev, ea_dev = eo_dev[lo:hi], ea_dev[lo:hi]
        # minibatch mode (constraint_net.py:304-317): one numpy permutation of min(N_nominal, N_expert) per backward
        # iteration that runs; all are drawn up front and the RNG is rewound to where the reference would have left it
        perm_dev, rng_states = None, None
        if self.batch_size is not None:
            if dp:
                raise NotImplementedError("cn_batch_size is not available in data-parallel mode")
            size = min(no.shape[0], eo_dev.shape[0])
            perms, rng_states = np.empty((max(int(iterations), 1), size), dtype=np.int32), [np.random.get_state()]
            for i in range(int(iterations)):
                perms[i] = np.random.permutation(size)
                rng_states.append(np.random.get_state())
            perm_dev = th.from_numpy(perms).to(self._dev)
        g = self.optimizer.param_groups[0]
        cfg = _lib.CnTrainCfg(
            iterations=int(iterations), importance_sampling=int(self.importance_sampling),
            per_step_is=int(self.per_step_importance_sampling), train_gail_lambda=int(bool(self.train_gail_lambda)),
            eps=float(self.eps), regularizer_coeff=float(self.regularizer_coeff or 0.0),
            target_kl_old_new=float(self.target_kl_old_new), target_kl_new_old=float(self.target_kl_new_old),
            lr=float(g["lr"]), adam_beta1=float(g["betas"][0]), adam_beta2=float(g["betas"][1]), adam_eps=float(g["eps"]),
            batch_size=0 if perm_dev is None else max(1, min(int(self.batch_size), perm_dev.shape[1])),
            perm=None if perm_dev is None else perm_dev.data_ptr())
        metrics = _lib.CnTrainMetrics()
        step = C.c_int64(self.optimizer.step_count)
        args = (C.byref(self._get_desc()), C.byref(cfg), _lib.ptr(no_dev), int(no.dtype == np.float64), _lib.ptr(na_dev),
                no.shape[0], _lib.ptr(off_dev), len(lengths), _lib.ptr(eo_dev), int(eo_dev.dtype == th.float64),
                _lib.ptr(ea_dev), eo_dev.shape[0], _lib.ptr(self._adam_m), _lib.ptr(self._adam_v), C.byref(step),
                C.byref(metrics))
        with th.cuda.device(self._dev):
            if not dp:
                _lib.check(_lib.lib().icrl_cn_train(*args, _lib.current_stream()))
            else:
                shapes = comm.shapes(no.shape[0], eo_dev.shape[0], len(lengths), self._dev)
                dist_desc = comm.descriptor(*shapes)
                _lib.check(_lib.lib().icrl_cn_train_dist(*args, C.byref(dist_desc), _lib.current_stream()))
                comm.advance(int(iterations))
        self.optimizer.step_count = int(step.value)
        if rng_states is not None:      # permutations are only drawn in iterations that get past the KL test
            np.random.set_state(rng_states[min(int(metrics.early_stop_itr), int(iterations))])
        keys = ("cn_loss", "expert_loss", "unweighted_nominal_loss", "nominal_loss", "regularizer_loss", "is_mean",
                "is_max", "is_min", "nominal_preds_max", "nominal_preds_min", "nominal_preds_mean", "expert_preds_max",
                "expert_preds_min", "expert_preds_mean")
        out = {f"backward/{k}": float(getattr(metrics, k)) for k in keys}
        if self.importance_sampling:
            out.update({"backward/kl_old_new": float(metrics.kl_old_new), "backward/kl_new_old": float(metrics.kl_new_old),
                        "backward/early_stop_itr": int(metrics.early_stop_itr)})
        return out

    def enable_data_parallel(self, comm=None, max_episodes: int = 65536) -> "ConstraintNet":
        """Data-parallel train() across the GPUs of one node (the reference has no distributed path): one
        process per GPU under torch.distributed, every rank passes the nominal episodes IT sampled to train(), the expert
        batch is split evenly by rank, and the partial sums of every backward iteration are exchanged inside the kernels
        over NVLink peer memory.  Parameters must start replicated; they stay bit-identical."""
        import torch.distributed as dist
        if comm is None:
            if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
                return self
            from .distributed import CnComm
            with th.cuda.device(self._dev):
                comm = CnComm(self._get_desc(), max_episodes)
        self.comm = comm
        return self

    # ---------------------------------------------------------------- checkpoints (constraint_net.py:323-402)
    def save(self, save_path):
        state_dict = dict(
            cn_network=OrderedDict((k, v.clone()) for k, v in self.network.state_dict().items()),
            cn_optimizer=self.optimizer.state_dict() if self.optimizer is not None else None,
            obs_dim=self.obs_dim, acs_dim=self.acs_dim, is_discrete=self.is_discrete,
            obs_select_dim=self.obs_select_dim, acs_select_dim=self.acs_select_dim, clip_obs=self.clip_obs,
            obs_mean=self.current_obs_mean, obs_var=self.current_obs_var, action_low=self.action_low,
            action_high=self.action_high, device=self.device, hidden_sizes=self.hidden_sizes)
        th.save(state_dict, save_path)

    @classmethod
    def load(cls, load_path: str, obs_dim=None, acs_dim=None, is_discrete=None, obs_select_dim=None,
             acs_select_dim=None, clip_obs=None, obs_mean=None, obs_var=None, action_low=None, action_high=None,
             device: str = "auto"):
        """Mirrors constraint_net.py:350-402 *including* its positional-argument shift: the loaded
        object has clip_obs = obs_mean (None => no obs clipping), current_obs_mean = obs_var, current_obs_var =
        action_low, action_low = action_high, action_high = None (=> no action clipping) and no optimiser, so that
        frozen `best_cn_model.pt` files produce the same costs as under the reference."""
        state_dict = th.load(load_path, weights_only=False)
        pick = lambda v, k: state_dict[k] if v is None else v
        obs_dim, acs_dim = pick(obs_dim, "obs_dim"), pick(acs_dim, "acs_dim")
        is_discrete = pick(is_discrete, "is_discrete")
        obs_select_dim, acs_select_dim = pick(obs_select_dim, "obs_select_dim"), pick(acs_select_dim, "acs_select_dim")
        clip_obs = pick(clip_obs, "clip_obs")
        obs_mean, obs_var = pick(obs_mean, "obs_mean"), pick(obs_var, "obs_var")
        action_low, action_high = pick(action_low, "action_low"), pick(action_high, "action_high")
        if device is None:
            device = state_dict["device"]
        hidden_sizes = state_dict["hidden_sizes"]
        # what the reference's 22 positional arguments actually bind to:
        cn = cls(obs_dim, acs_dim, hidden_sizes, None, None, None, None, is_discrete,
                 regularizer_coeff=None, obs_select_dim=obs_select_dim, acs_select_dim=acs_select_dim,
                 optimizer_class=None, optimizer_kwargs=None, no_importance_sampling=None,
                 per_step_importance_sampling=clip_obs, clip_obs=obs_mean, initial_obs_mean=obs_var,
                 initial_obs_var=action_low, action_low=action_high, action_high=None,
                 target_kl_old_new=None, target_kl_new_old=device)
        cn.load_network_state_dict(state_dict["cn_network"])
        return cn
