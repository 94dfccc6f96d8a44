for S in 68 72 76 64; do
  python bench.py --steps 20 --warmup 5 --stage-sms $S --no-e2e --no-cpu-baseline > gpurun_out/v4d_$S.json 2> gpurun_out/v4d_$S.err
done
python bench.py --steps 20 --warmup 5 --config 4 --stage-sms 72 --no-e2e --no-cpu-baseline > gpurun_out/v4d_c4_72.json 2> gpurun_out/v4d_c4_72.err
python bench.py --steps 20 --warmup 5 --config 4 --stage-sms 90 --no-e2e --no-cpu-baseline > gpurun_out/v4d_c4_90.json 2> gpurun_out/v4d_c4_90.err
python bench.py --steps 20 --warmup 5 --config 0 --stage-sms 48 --no-e2e --no-cpu-baseline > gpurun_out/v4d_c0_48.json 2> gpurun_out/v4d_c0_48.err
python bench.py --steps 20 --warmup 5 --config 0 --stage-sms 32 --no-e2e --no-cpu-baseline > gpurun_out/v4d_c0_32.json 2> gpurun_out/v4d_c0_32.err
