set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "field_fast or end_to_end_default or reference_golden" 2>&1 | tail -3
for m in A B; do
for v in _base fenerf_b200 _lag1 _lag3 _lag4 _base fenerf_b200; do
  printf "%-12s %s " $v $m
  FENERF_B200_LIB=fenerf_b200/lib$v.so timeout 120 python tools/ab_field.py $m 12 2>/dev/null | sed 's/.*median/median/'
done; done | tee gpurun_out/r02_seqlag.txt
