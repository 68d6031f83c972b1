"""cProfile of the host-API iteration (bench.py's e2e leg): where the host-side milliseconds go."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402

import bench  # noqa: E402
from icrl_b200.learner import WORKLOADS, DeviceLearner  # noqa: E402

w = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "halfcheetah"]
dl = DeviceLearner(w, seed=0)
hl = bench.HostLearner(w, dl)
for _ in range(2):
    hl.run()
th.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
hl.run()
th.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
