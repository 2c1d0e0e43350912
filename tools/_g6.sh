cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/quick_toon_bench.py 0 A 16384 32768 B 16384 32768 2>&1 | tail -3
timeout 600 python tools/hook_bench.py 1024 > gpurun_out/hook_bench.json 2> gpurun_out/hook_bench.err; tail -c 3000 gpurun_out/hook_bench.json; tail -5 gpurun_out/hook_bench.err
