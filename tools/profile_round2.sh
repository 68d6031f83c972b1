#!/bin/bash
# Round-2 one-GPU evidence run for profiles/: tests, the default bench line, ncu launch list of one timed iteration,
# ncu --set full captures of the kernels this round changed (K4 single cluster, many-cluster K4, K3, K2 gradient kernel),
# compute-sanitizer over the smoke path.
# Usage (under gpurun, from the repo root):  bash tools/profile_round2.sh r02 [tests bench launches ncu sanitize]
set -u
R=${1:-r02}
shift || true
WHAT=${*:-tests bench launches ncu sanitize}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=.
for what in $WHAT; do
case $what in
tests)
    timeout 1200 python -m pytest tests -q -m gpu --timeout 300 > $OUT/pytest_gpu_$R.txt 2>&1
    echo "pytest rc $?" >> $OUT/pytest_gpu_$R.txt
    tail -4 $OUT/pytest_gpu_$R.txt ;;
bench)
    timeout 900 python bench.py > $OUT/bench_hc_$R.json 2> $OUT/bench_hc_$R.err
    echo "bench rc $?"; tail -3 $OUT/bench_hc_$R.err; head -c 600 $OUT/bench_hc_$R.json; echo ;;
launches)
    # launch list of one timed iteration (cold-cache, serialised: shares, not absolutes)
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_$R.csv \
        python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-workloads --no-sweep > $OUT/bench_under_ncu_$R.log 2>&1
    echo "launches rc $?"; wc -l $OUT/launches_$R.csv ;;
ncu)
    K4_STEPS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ppo_train_kernel -s 1 -c 1 -f -o $OUT/k4_$R \
        python tools/profile_target.py k4 > $OUT/ncu_k4_$R.log 2>&1
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:ppo_train_kernel -s 2 -c 1 -f -o $OUT/k4wide_$R \
        python tools/k4_wide_time.py antwall 1048576 13107 2 > $OUT/ncu_k4wide_$R.log 2>&1
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:dual_gae_kernel -s 2 -c 4 -f -o $OUT/k3_$R \
        python tools/profile_target.py k3 > $OUT/ncu_k3_$R.log 2>&1
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:cn_forward -s 2 -c 4 -f -o $OUT/k1_$R \
        python tools/profile_target.py k1 > $OUT/ncu_k1_$R.log 2>&1
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:cn_grad_kernel -s 1 -c 1 -f -o $OUT/k2_$R \
        python tools/profile_target.py k2 > $OUT/ncu_k2_$R.log 2>&1
    ls -la $OUT/*_$R.ncu-rep ;;
sanitize)
    bash tools/sanitize.sh $R ;;
esac
done
