cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_toon_tp_gpu.py tests/test_toon_gpu.py -x -q 2>&1 | grep -v "^ERROR\|^WARNING" | tail -3
timeout 300 python tools/quick_toon_bench.py 0 A 16384 32768 B 16384 32768 P 16384 32768 A 2048 131072 M 0 4096 2>&1 | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_chain.json 2> gpurun_out/bench_chain.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_chain.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e_cabi']['value'], d['stages'], d['roofline']['frac'])
PY
