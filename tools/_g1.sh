set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_toon_tp_gpu.py tests/test_toon_gpu.py -x -q 2>&1 | tail -15
timeout 300 python tools/quick_toon_bench.py 0 2>&1 | tail -5
timeout 300 python tools/quick_toon_bench.py 8 2>&1 | tail -5
timeout 300 python tools/quick_toon_bench.py 0 C 16384 32768 M 0 4096 A 262144 1 A 16384 1 2>&1 | tail -6
