run() { python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline "$@"; }
run --config 1 > gpurun_out/v4f_c1_auto.json 2> gpurun_out/v4f_c1_auto.err
run --config 1 --stage-sms 52 > gpurun_out/v4f_c1_52.json 2>/dev/null
run --config 1 --stage-sms 64 > gpurun_out/v4f_c1_64.json 2>/dev/null
run --config 1 --stage-sms 48 > gpurun_out/v4f_c1_48.json 2>/dev/null
run --config 0 > gpurun_out/v4f_c0_auto.json 2> gpurun_out/v4f_c0_auto.err
run --config 0 --stage-sms 32 > gpurun_out/v4f_c0_32.json 2>/dev/null
run --config 0 --stage-sms 72 > gpurun_out/v4f_c0_72.json 2>/dev/null
run --config 4 > gpurun_out/v4f_c4_auto.json 2> gpurun_out/v4f_c4_auto.err
run --config 4 --stage-sms 56 > gpurun_out/v4f_c4_56.json 2>/dev/null
run --config 4 --stage-sms 64 > gpurun_out/v4f_c4_64.json 2>/dev/null
timeout 900 python -m pytest tests -m gpu -q -x -k "cov or full or fused or pipeline or hooks" > gpurun_out/v4f_tests.log 2>&1; tail -2 gpurun_out/v4f_tests.log
