set -u
mkdir -p gpurun_out
export PYTHONPATH=.
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_dp_gpu.py -q -m gpu --timeout 400 -k "4-auto-hc or 4-rsag-hc or oracle[4 or 2-auto-hc or 2-auto-ant" > gpurun_out/pytest_dp4_r02c.txt 2>&1
tail -4 gpurun_out/pytest_dp4_r02c.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_dp4_r02c.json 2> gpurun_out/bench_dp4_r02c.err
echo "bench dp4 rc $?"; tail -2 gpurun_out/bench_dp4_r02c.err; grep -o '"value": [0-9.]*' gpurun_out/bench_dp4_r02c.json | head -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_dp2_r02c.json 2> gpurun_out/bench_dp2_r02c.err
echo "bench dp2 rc $?"; tail -2 gpurun_out/bench_dp2_r02c.err; grep -o '"value": [0-9.]*' gpurun_out/bench_dp2_r02c.json | head -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 tools/dp_k4_time.py halfcheetah 2>&1 | grep -E "world|ppo timing" | grep -v "cta [135]" | tail -3
