set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_k2_gpu.py -q -m gpu --timeout 120 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_dp2_r02a.json 2> gpurun_out/bench_dp2_r02a.err
echo "bench dp2 rc $?"; tail -15 gpurun_out/bench_dp2_r02a.err; head -c 1500 gpurun_out/bench_dp2_r02a.json
timeout 900 python -m pytest tests/test_dp_gpu.py -q -m gpu --timeout 300 -k "2-direct or 2-rsag or wide or sharded" > gpurun_out/pytest_dp2_r02a.txt 2>&1
tail -30 gpurun_out/pytest_dp2_r02a.txt
