set -u
mkdir -p gpurun_out
echo "== K3 variants"; python tools/kernel_times.py k3 4194304 16777216; ICRL_K3_NT=128 python tools/kernel_times.py k3 4194304 16777216
echo "== K4 wide timing"
ICRL_PPO_TIMING=1 python tools/k4_wide_time.py antwall 1048576 2>&1 | tail -8
ICRL_PPO_TIMING=1 python tools/k4_wide_time.py halfcheetah 1048576 2>&1 | tail -8
echo "== single cluster timing (Ant, HC)"
ICRL_PPO_TIMING=1 python tools/k4_wide_time.py antwall 10240 128 2 2>&1 | tail -5
