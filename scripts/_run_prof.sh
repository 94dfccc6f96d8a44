python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02b_cfg1.json 2> gpurun_out/bench_r02b_cfg1.err || exit 1
tail -c 400 gpurun_out/bench_r02b_cfg1.json
timeout 900 ncu --set full --clock-control none -k "regex:stage_group|stage_tma|contraction_tc|autocorr_tc" -c 10 -o gpurun_out/prof_r02b -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r02b.log 2>&1
tail -2 gpurun_out/ncu_r02b.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:nsgp|stage_|contraction_|autocorr_|sgd_|repre_|class_index|proj_|split|segment|threshold|normalize|greedy" -c 400 --csv --log-file gpurun_out/launches_r02b.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r02b_list.log 2>&1
tail -2 gpurun_out/ncu_r02b_list.log; wc -l gpurun_out/launches_r02b.csv
