cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_manager_gpu.py tests/test_toon_tp_gpu.py -x -q -m gpu 2>&1 | grep -v "^ERROR\|^WARNING" | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_chain.json 2> gpurun_out/bench_chain.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_chain.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e_cabi']['value'], d['stages'], d['roofline']['frac'])
PY
tail -3 gpurun_out/bench_chain.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
