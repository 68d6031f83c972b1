"""TEST INFRASTRUCTURE -- CPU restatement ("oracle") of the reference's ICRL learner hot path.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import this package.  The product (`icrl_b200/`) never does:
it calls the sm_100a CUDA library through the C-ABI and fails loudly without it.

Parity status: the reference (shehryar-malik/icrl) ships NO tests or golden vectors
for this path ("parity unpinned" by the reference's own suite).  The oracle is
therefore pinned against outputs of the *unmodified reference executed in the build
container* (`tests/golden/make_golden.py` -> `tests/golden/*.npz`, via
`oracle/ref_shim.py`), and `tests/test_oracle_vs_golden.py` checks every function
here against those fixtures.

The reference's arithmetic for this path lives in two third-party libraries that are
not vendored under /root/reference: torch (pinned 1.5.0 in README.md:11; 2.11.0 here)
and numpy (pinned 1.17.5; 2.3.5 here).  The oracle restates the algorithm on top of
the same two libraries' CPU kernels (explicit weight tensors + autograd, no
nn.Module / torch.distributions / optim objects), so that it is float-for-float the
same arithmetic while being structurally independent of the reference's classes.

Modules:
  cn.py    constraint net: input preparation, forward/cost (K1), IS weights + train step (K2)
  gae.py   dual reward/cost GAE (K3)
  ppo.py   ActorTwoCritics forward, PPO-Lagrangian minibatch update, dual step (K4)
"""
