cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc; nvidia-smi -L | wc -l
run() { # n workload port steps
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $1 --steps $4 --warmup 3 --workload $2 > gpurun_out/r02_bench_$2_n$1.json 2> gpurun_out/r02_bench_$2_n$1.err
  tail -c 300 gpurun_out/r02_bench_$2_n$1.json | cut -c1-300; tail -2 gpurun_out/r02_bench_$2_n$1.err | cut -c1-300
}
run 8 chain 29521 5
run 8 scan 29522 20
run 4 chain 29523 5
run 4 scan 29524 20
run 2 scan 29525 20
timeout 600 python bench.py --workload scan --steps 20 --warmup 3 > gpurun_out/r02_bench_scan_n1.json 2> gpurun_out/r02_bench_scan_n1.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_chain_n1.json 2> gpurun_out/r02_bench_chain_n1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*_n*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d.get('roofline',{}).get('frac'))
    except Exception as e:
        print(f, "ERR", e)
PY
