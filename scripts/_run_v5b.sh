run() { python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline "$@"; }
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/v5_tests.log 2>&1; tail -2 gpurun_out/v5_tests.log
run --config 1 > gpurun_out/v5b_c1_auto.json 2> gpurun_out/v5b_c1_auto.err
for S in 40 44 48 52 60 64; do run --config 1 --stage-sms $S > gpurun_out/v5b_c1_$S.json 2>/dev/null; done
run --config 0 > gpurun_out/v5b_c0_auto.json 2>/dev/null
run --config 4 > gpurun_out/v5b_c4_auto.json 2>/dev/null
run --config 4 --stage-sms 72 > gpurun_out/v5b_c4_72.json 2>/dev/null
run --config 4 --stage-sms 56 > gpurun_out/v5b_c4_56.json 2>/dev/null
