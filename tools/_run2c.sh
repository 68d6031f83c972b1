set -u
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_dp_gpu.py -q -m gpu --timeout 300 -k "2-auto" 2>&1 | tail -3
for wl in halfcheetah antwall; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_k4_time.py $wl 2>&1 | grep -E "world|ppo timing" | grep -v "cta [35]"
done
