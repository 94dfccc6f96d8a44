run() { python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline "$@"; }
for S in 8056 8052 12056 6056 4056 8060; do run --config 1 --stage-sms $S > gpurun_out/res_$S.json 2>/dev/null; done
