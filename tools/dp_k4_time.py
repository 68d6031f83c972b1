"""Data-parallel K4 under torchrun: launch time per optimiser step and (rank 0) the per-phase cycles incl. the exchange split.
Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dp_k4_time.py [workload]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402
import torch.distributed as dist  # noqa: E402

from icrl_b200.distributed import PpoComm  # noqa: E402
from icrl_b200.learner import WORKLOADS, DeviceLearner  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
th.cuda.set_device(local)
dist.init_process_group("nccl", device_id=th.device("cuda", local))
w = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "halfcheetah"]
w = type(w)(**{**w.__dict__, "rollouts": 3, "backward_iters": 0})
comm = PpoComm() if world > 1 else None
learner = DeviceLearner(w, seed=100 + rank, device=th.device("cuda", local), comm=comm, param_seed=0)
for rep in range(3):
    if rep == 2 and rank == 0:
        os.environ["ICRL_PPO_TIMING"] = "1"
    dist.barrier(); th.cuda.synchronize()
    s, e = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    s.record(); learner.run(); e.record(); th.cuda.synchronize()
    learner.check()
    steps = w.rollouts * learner.steps_taken_per_rollout()
    if rank == 0:
        print(f"world {world} {sys.argv[1] if len(sys.argv) > 1 else 'halfcheetah'} rep {rep}: {s.elapsed_time(e):.2f} ms for {steps} steps = "
              f"{s.elapsed_time(e) * 1e3 / steps:.2f} us/step (incl. K1/K5/K3 per rollout)", flush=True)
dist.destroy_process_group()
