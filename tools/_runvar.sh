set -u
export PYTHONPATH=.
cp icrl_b200/libicrl_b200.so /tmp/orig.so
for v in prev cur; do
  cp tools/_variants/$v.so icrl_b200/libicrl_b200.so
  for wl in halfcheetah antwall lapgrid pointcircle; do
    VAR=$v K4_STEPS=0 python - $wl <<'PY'
import os, sys
sys.argv = ["x", "k4", sys.argv[1]]
import torch as th
exec(open("tools/profile_target.py").read().split("th.cuda.synchronize()\nif what")[0])
w2 = type(w)(**{**w.__dict__, "backward_iters": 0})
learner.w = w2
learner.max_steps = 0
ts = []
for _ in range(4):
    s, e = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    th.cuda.synchronize(); s.record(); learner.run(); e.record(); th.cuda.synchronize()
    ts.append(s.elapsed_time(e))
steps = 2 * learner.steps_taken_per_rollout()
print(f"{os.environ.get('VAR')} {wname}: {min(ts[1:]):.2f} ms / {steps} steps = {min(ts[1:]) * 1e3 / steps:.3f} us/step (2 rollouts incl. K1/K5/K3)")
PY
  done
done
cp /tmp/orig.so icrl_b200/libicrl_b200.so
