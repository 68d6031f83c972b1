set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_k3_gpu.py -q -m gpu --timeout 120 2>&1 | tail -8
timeout 300 python tools/kernel_times.py k1,k3,k5 2>&1 | tee gpurun_out/kernel_times_r02a.txt
