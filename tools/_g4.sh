cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_toon_tp_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tools/quick_toon_bench.py 0 A 16384 32768 A 2048 131072 B 16384 32768 M 0 4096 A 262144 1 A 16384 1 2>&1 | tail -7
timeout 600 ncu --set full --clock-control none --import-source on -k regex:toon_tp -s 2 -c 1 -o gpurun_out/tp_a16k -f python tools/quick_toon_bench.py 0 A 16384 8192 > gpurun_out/ncu_tp.log 2>&1
tail -2 gpurun_out/ncu_tp.log
