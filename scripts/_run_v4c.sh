for K in 0 16 32 64 48; do
  echo "skip=$K"
  NSGP_STAGE_SKIP=$K NSGP_BRINGUP_LIB=1 NSGP_TIMELINE=1 NOPROF=1 TL_N=4 STAGE_SMS=0 timeout 300 python scripts/bench_cov.py 8 deferred 2>&1 | tail -4
done
