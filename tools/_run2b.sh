set -u
mkdir -p gpurun_out
export PYTHONPATH=.
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_k4_gpu.py tests/test_drivers_gpu.py tests/test_k5_gpu.py -q -m gpu --timeout 300 2>&1 | tail -3
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --no-sweep --no-workloads --no-cpu > gpurun_out/bench_1gpu_r02c.json 2> gpurun_out/bench_1gpu_r02c.err
echo "bench 1 rc $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_1gpu_r02c.json').read().strip().split('\n')[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
PY
timeout 900 python -m pytest tests/test_dp_gpu.py -q -m gpu --timeout 300 -k "2-auto or 2-direct-hc or 2-rsag-hc" 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-sweep --no-workloads > gpurun_out/bench_dp2_r02c.json 2> gpurun_out/bench_dp2_r02c.err
echo "bench dp2 rc $?"; tail -3 gpurun_out/bench_dp2_r02c.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_dp2_r02c.json').read().strip().split('\n')[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['dp_parity']['ok'])
PY
