"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's online cost normalisation, applied to a whole
rollout at once.

Follows, per time step t of a [T, E] rollout (numpy float64 state exactly as the reference keeps it):
    stable_baselines3/common/vec_env/vec_normalize.py:232-257   VecNormalizeWithCost.step_wait / _update_cost /
                                                                normalize_cost  (+ :276-282 reset)
    stable_baselines3/common/running_mean_std.py:19-39          RunningMeanStd.update / update_from_moments

Pinned against tests/golden/venv_{all,nocost}.npz, which were produced by the reference's own wrapper classes
(tests/golden/make_golden.py::golden_venv).  Only tests/, smoke() and bench.py's cpu_baseline may import this.
"""
import numpy as np


def initial_state(E, rms_epsilon=1e-4):
    """State right after VecNormalizeWithCost.__init__: rms (mean 0, var 1, count 1e-4), cost_ret zeros."""
    return dict(mean=np.float64(0.0), var=np.float64(1.0), count=np.float64(rms_epsilon), cost_ret=np.zeros(E))


def rms_update(state, batch):
    """running_mean_std.py:19-39 on a 1-D batch."""
    batch_mean, batch_var, batch_count = np.mean(batch, axis=0), np.var(batch, axis=0), batch.shape[0]
    delta = batch_mean - state["mean"]
    tot_count = state["count"] + batch_count
    new_mean = state["mean"] + delta * batch_count / tot_count
    m_a = state["var"] * state["count"]
    m_b = batch_var * batch_count
    m_2 = m_a + m_b + np.square(delta) * state["count"] * batch_count / (state["count"] + batch_count)
    state["mean"], state["var"], state["count"] = new_mean, m_2 / (state["count"] + batch_count), batch_count + state["count"]


def reset(state, training=True):
    """vec_normalize.py:276-282: cost_ret := 0 and (when training) one statistics update with that zero batch."""
    state["cost_ret"] = np.zeros_like(state["cost_ret"])
    if training:
        rms_update(state, state["cost_ret"])
    return state


def normalize_rollout(orig_costs, news, state, cost_gamma=0.99, epsilon=1e-8, clip_cost=10.0, norm_cost=True,
                      training=True):
    """orig_costs [T, E] (what the cost function returned at each step), news [T, E] bool (episode ended at that
    step).  Returns the float32 costs the rollout buffer would have stored; `state` is advanced in place."""
    T, E = orig_costs.shape
    out = np.empty((T, E), np.float32)
    for t in range(T):
        cost = np.asarray(orig_costs[t])
        if training:
            state["cost_ret"] = state["cost_ret"] * cost_gamma + cost
            rms_update(state, state["cost_ret"])
        c = cost
        if norm_cost:
            c = np.clip(cost / np.sqrt(state["var"] + epsilon), -clip_cost, clip_cost)
        out[t] = c
        state["cost_ret"][np.asarray(news[t], bool)] = 0
    return out
