set -x
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/v4_tests.log 2>&1; echo "tests rc=$?" 
tail -5 gpurun_out/v4_tests.log
for S in 40 48 56 64 72; do
  NSGP_BRINGUP_LIB=1 NSGP_TIMELINE=1 NOPROF=1 TL_N=8 STAGE_SMS=$S timeout 300 python scripts/bench_cov.py 8 deferred > gpurun_out/v4_cov_$S.log 2>&1
  grep "join_each=0" gpurun_out/v4_cov_$S.log
done
NSGP_STAGE_TMA=0 NSGP_BRINGUP_LIB=1 NSGP_TIMELINE=1 NOPROF=1 TL_N=8 STAGE_SMS=64 timeout 300 python scripts/bench_cov.py 8 deferred > gpurun_out/v4_cov_reg64.log 2>&1
grep "join_each=0" gpurun_out/v4_cov_reg64.log
NSGP_BRINGUP_LIB=1 NSGP_TIMELINE=1 NOPROF=1 TL_N=8 STAGE_SMS=0 timeout 300 python scripts/bench_cov.py 8 deferred > gpurun_out/v4_cov_0.log 2>&1
grep "join_each=0" gpurun_out/v4_cov_0.log
