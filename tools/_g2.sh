cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:toon_tp -s 2 -c 1 -o gpurun_out/tp_a16k -f python tools/quick_toon_bench.py 0 A 16384 8192 > gpurun_out/ncu_tp.log 2>&1
tail -3 gpurun_out/ncu_tp.log
