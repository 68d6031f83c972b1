"""Entry point with the reference's dispatch convention (run_me.py:1-32): `python run_me.py <file_to_run> [flags]`,
every driver ignoring the first positional argument.  Only the ICRL learner path is built here: `icrl`, `cpg` and
the `run_policy` tool that writes expert rollouts; the other reference baselines (gail, airl, pruning) are out of
scope."""
import sys

if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit("usage: python run_me.py {icrl,cpg,run_policy} [flags]")
    file_to_run = sys.argv[1]
    if file_to_run == "cpg":
        from icrl_b200.cpg import main
    elif file_to_run == "icrl":
        from icrl_b200.icrl import main
    elif file_to_run == "run_policy":
        from icrl_b200.run_policy import main
    elif file_to_run in ("gail", "airl", "random_agent", "pruning/train.py", "pruning/pruning_env.py"):
        raise NotImplementedError("File %s is outside the ICRL learner hot path and not built here" % file_to_run)
    else:
        raise ValueError("File %s not defined" % file_to_run)
    main()
