"""Entry point with the reference's dispatch convention: `python run_me.py <file_to_run> [flags]`, every driver ignoring
the first positional argument (reference run_me.py:1-32).  Only the ICRL learner path is built here: `icrl`, `cpg` and the
`run_policy` tool that writes expert rollouts; the reference's other baselines are out of scope."""
import importlib
import sys

DRIVERS = {"icrl": "icrl_b200.icrl", "cpg": "icrl_b200.cpg", "run_policy": "icrl_b200.run_policy"}
NOT_BUILT = ("gail", "airl", "random_agent", "pruning/train.py", "pruning/pruning_env.py")

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else None
    if which in DRIVERS:
        importlib.import_module(DRIVERS[which]).main()
    elif which in NOT_BUILT:
        raise NotImplementedError("File %s is outside the ICRL learner hot path and not built here" % which)
    elif which is None:
        raise SystemExit("usage: python run_me.py {%s} [flags]" % ",".join(DRIVERS))
    else:
        raise ValueError("File %s not defined" % which)
