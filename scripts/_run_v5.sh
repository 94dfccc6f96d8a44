timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/v5_tests.log 2>&1; tail -2 gpurun_out/v5_tests.log
for S in 36 40 44 48 56 64; do
  echo "S=$S"
  NSGP_BRINGUP_LIB=1 NSGP_TIMELINE=1 NOPROF=1 TL_N=8 STAGE_SMS=$S timeout 300 python scripts/bench_cov.py 8 deferred 2>&1 | tail -9 | egrep "join_each=0|gram-autocorr|stage" | head -4
done
echo "S=0"; NSGP_BRINGUP_LIB=1 NSGP_TIMELINE=1 NOPROF=1 TL_N=4 STAGE_SMS=0 timeout 300 python scripts/bench_cov.py 8 deferred 2>&1 | tail -4
for B in 16 2; do for S in 40 72; do echo "B=$B S=$S"; NSGP_BRINGUP_LIB=1 NSGP_TIMELINE=1 NOPROF=1 TL_N=8 STAGE_SMS=$S timeout 300 python scripts/bench_cov.py $B deferred 2>&1 | tail -8 | grep -A1 "gram-autocorr" | head -2; done; done
