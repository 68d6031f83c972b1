set -u
python tools/kernel_times.py k3 1048576 4194304 16777216
echo "== CG=4"; ICRL_K3_CG=4 python tools/kernel_times.py k3 1048576 4194304 16777216
echo "== NT=256"; ICRL_K3_NT=256 python tools/kernel_times.py k3 16777216
python -m pytest tests/test_k3_gpu.py -q -m gpu 2>&1 | tail -2
