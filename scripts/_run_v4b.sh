for S in 56 60 64; do
  NSGP_BRINGUP_LIB=1 NSGP_TIMELINE=1 python bench.py --steps 20 --warmup 5 --stage-sms $S --no-e2e --no-cpu-baseline > gpurun_out/v4b_$S.json 2> gpurun_out/v4b_$S.err
  python bench.py --steps 20 --warmup 5 --stage-sms $S --no-e2e --no-cpu-baseline > gpurun_out/v4c_$S.json 2> gpurun_out/v4c_$S.err
  NSGP_SIDE_PRIO=0 python bench.py --steps 20 --warmup 5 --stage-sms $S --no-e2e --no-cpu-baseline > gpurun_out/v4p_$S.json 2> gpurun_out/v4p_$S.err
done
