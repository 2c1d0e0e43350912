cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^ERROR\|^WARNING" | tail -8
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_chain.json 2> gpurun_out/bench_chain.err; tail -c 4000 gpurun_out/bench_chain.json; tail -5 gpurun_out/bench_chain.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_chain_ref.json 2>&1; tail -c 600 gpurun_out/bench_chain_ref.json
