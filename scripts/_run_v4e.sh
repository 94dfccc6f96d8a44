for B in 16 2; do
for S in 40 56 72 88; do
  echo "B=$B S=$S"
  NSGP_BRINGUP_LIB=1 NSGP_TIMELINE=1 NOPROF=1 TL_N=8 STAGE_SMS=$S timeout 300 python scripts/bench_cov.py $B deferred 2>&1 | tail -8 | grep -A1 "gram-autocorr" | head -2
done
echo "B=$B S=0"
NSGP_BRINGUP_LIB=1 NSGP_TIMELINE=1 NOPROF=1 TL_N=4 STAGE_SMS=0 timeout 300 python scripts/bench_cov.py $B deferred 2>&1 | tail -4
done
