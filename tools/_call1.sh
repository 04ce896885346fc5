set -x
FENERF_B200_LIB=fenerf_b200/lib_share.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "field_fast" 2>&1 | tail -3
for m in A B; do
for v in fenerf_b200 _share _share16 _lag1 _lag2 _lag3; do
  printf "%-12s %s " $v $m
  FENERF_B200_LIB=fenerf_b200/lib$v.so timeout 100 python tools/ab_field.py $m 8 2>/dev/null | sed 's/.*median/median/'
  echo
done; done | tee gpurun_out/r02_seqlag.txt
