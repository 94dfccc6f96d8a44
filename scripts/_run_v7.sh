timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
run() { python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline "$@"; }
run --config 1 > gpurun_out/v8_c1_auto.json 2> gpurun_out/v8_c1_auto.err
for S in 56 64; do run --config 1 --stage-sms $S > gpurun_out/v8_c1_$S.json 2>/dev/null; done
NSGP_BRINGUP_LIB=1 NSGP_TIMELINE=1 run --config 1 > gpurun_out/v8_c1_tl.json 2> gpurun_out/v8_c1_tl.err
