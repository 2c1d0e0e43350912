cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:toon_tp -s 2 -c 1 -o gpurun_out/tp_b16k -f python tools/quick_toon_bench.py 0 B 16384 4096 > gpurun_out/ncu_tpb.log 2>&1
tail -1 gpurun_out/ncu_tpb.log
