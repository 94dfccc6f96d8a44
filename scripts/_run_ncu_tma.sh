STAGE_SMS=40 NOPROF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_tma -s 10 -c 1 -o gpurun_out/tma40 -f python scripts/bench_cov.py 8 deferred > gpurun_out/ncu_tma40.log 2>&1
tail -3 gpurun_out/ncu_tma40.log
