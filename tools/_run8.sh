set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1500 python -m pytest tests/test_dp_gpu.py -q -m gpu --timeout 400 -k "4-auto-hc or 4-rsag-hc or 8-auto or wide_ppo_matches_oracle[4 or wide_ppo_matches_oracle[8 or oracle[4-cn or oracle[8-cn" > gpurun_out/pytest_dp48_r02.txt 2>&1
tail -15 gpurun_out/pytest_dp48_r02.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_dp8_r02.json 2> gpurun_out/bench_dp8_r02.err
echo "bench dp8 rc $?"; tail -5 gpurun_out/bench_dp8_r02.err; grep -o '"value": [0-9.]*' gpurun_out/bench_dp8_r02.json | head -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_dp4_r02.json 2> gpurun_out/bench_dp4_r02.err
echo "bench dp4 rc $?"; tail -5 gpurun_out/bench_dp4_r02.err; grep -o '"value": [0-9.]*' gpurun_out/bench_dp4_r02.json | head -3
