cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_toon_tp_gpu.py tests/test_manager_gpu.py -x -q 2>&1 | grep -v "^ERROR\|^WARNING" | tail -3
timeout 300 python tools/quick_toon_bench.py 0 A 16384 32768 B 16384 32768 P 16384 32768 N 16384 32768 2>&1 | tail -4
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_chain.json 2> gpurun_out/bench_chain.err; tail -c 3300 gpurun_out/bench_chain.json; tail -3 gpurun_out/bench_chain.err
