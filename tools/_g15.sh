cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_toon_tp_gpu.py tests/test_toon_gpu.py tests/test_manager_gpu.py -x -q 2>&1 | grep -v "^ERROR\|^WARNING" | tail -3
timeout 200 python tools/quick_toon_bench.py 0 A 16384 32768 P 16384 32768 M 0 4096 A 262144 1 2>&1 | tail -4
