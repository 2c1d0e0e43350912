cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_chain_n2.json 2> gpurun_out/bench_chain_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_chain_n2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['e2e_cabi']['value'], d['stages']['toon_stage'])
except Exception as e:
    print("ERR", e); print(open('gpurun_out/bench_chain_n2.err').read()[-3000:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --workload scan > gpurun_out/bench_scan_n2.json 2> gpurun_out/bench_scan_n2.err
tail -c 400 gpurun_out/bench_scan_n2.json; tail -3 gpurun_out/bench_scan_n2.err
