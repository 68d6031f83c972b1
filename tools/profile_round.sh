#!/bin/bash
# One-GPU evidence run for profiles/: tests, bench lines, ncu launch list, ncu --set full captures.
# Usage (under gpurun, from the repo root):  bash tools/profile_round.sh r01
set -u
R=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=.
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > $OUT/pytest_gpu_$R.txt
python bench.py > $OUT/bench_hc_$R.json 2> $OUT/bench_hc_$R.err
python bench.py --workload antwall --no-cpu > $OUT/bench_ant_$R.json 2> $OUT/bench_ant_$R.err
python bench.py --workload lapgrid --no-cpu > $OUT/bench_lgw_$R.json 2> $OUT/bench_lgw_$R.err
python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_hc_ref_$R.json 2> $OUT/bench_hc_ref_$R.err
# launch list of one timed iteration (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches_$R.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > $OUT/bench_under_ncu_$R.log 2>&1
# full captures: K4 (320 optimiser steps), K1 / K3 at rollout size and at 4M rows, K5
K4_STEPS=${K4_STEPS:-0} ncu --set full --clock-control none --import-source on -k regex:ppo_train_kernel -s 1 -c 1 -f -o $OUT/k4_$R \
    python tools/profile_target.py k4 > $OUT/ncu_k4_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cn_forward_kernel -s 2 -c 4 -f -o $OUT/k1_$R \
    python tools/profile_target.py k1 > $OUT/ncu_k1_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dual_gae_kernel -s 2 -c 4 -f -o $OUT/k3_$R \
    python tools/profile_target.py k3 > $OUT/ncu_k3_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cost_norm -s 1 -c 1 -f -o $OUT/k5_$R \
    python tools/profile_target.py iter > $OUT/ncu_k5_$R.log 2>&1
ls -la $OUT | tail -30
