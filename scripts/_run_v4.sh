set -x
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/v4_tests.log 2>&1; echo "tests rc=$?" 
tail -3 gpurun_out/v4_tests.log
for S in 44 52 56 60 64 68; do
  NSGP_BRINGUP_LIB=1 NSGP_TIMELINE=1 NOPROF=1 TL_N=8 STAGE_SMS=$S timeout 300 python scripts/bench_cov.py 8 deferred > gpurun_out/v4_cov_$S.log 2>&1
  grep "join_each=0" gpurun_out/v4_cov_$S.log
done
NSGP_BRINGUP_LIB=1 NSGP_TIMELINE=1 NOPROF=1 TL_N=8 STAGE_SMS=0 timeout 300 python scripts/bench_cov.py 8 deferred > gpurun_out/v4_cov_0.log 2>&1
grep "join_each=0" gpurun_out/v4_cov_0.log
python bench.py --steps 20 --warmup 5 --stage-sms 56 > gpurun_out/v4_bench56.json 2> gpurun_out/v4_bench56.err
python bench.py --steps 20 --warmup 5 --stage-sms 64 --no-e2e --no-cpu-baseline > gpurun_out/v4_bench64.json 2> gpurun_out/v4_bench64.err
python bench.py --steps 20 --warmup 5 --stage-sms 48 --no-e2e --no-cpu-baseline > gpurun_out/v4_bench48.json 2> gpurun_out/v4_bench48.err
