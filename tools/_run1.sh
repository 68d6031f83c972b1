set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -k "not dp" > gpurun_out/pytest_gpu_r02a.txt 2>&1
echo "pytest rc $?" >> gpurun_out/pytest_gpu_r02a.txt
grep -E "^FAILED|^ERROR|passed|failed|k4x " gpurun_out/pytest_gpu_r02a.txt | head -60
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err
echo "bench rc $?"; tail -5 gpurun_out/bench_r02a.err; head -c 3000 gpurun_out/bench_r02a.json
