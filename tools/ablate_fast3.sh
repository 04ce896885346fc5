#!/bin/bash
# Ablation timings of the tcgen05 point-network launch (debug build, WRONG results by design):
#   FENERF_NVCC_DEFINES=-DFENERF_DEBUG_SHORT_LOADS FENERF_B200_LIB=fenerf_b200/lib_dbg.so python -m fenerf_b200.build
#   bash tools/ablate_fast3.sh [A|B]
# mask bits: 1 short weight loads, 2 FiLM epilogue hand-offs only, 4 no tcgen05.mma, 8 no sin, 16 no activation stores, 32 no tcgen05.ld
m=${1:-A}
for mask in 0 1 2 3 4 5 6 7 8 16 32 40 48 56 9 36; do
  printf "mask %2d: " $mask
  FENERF_B200_DEBUG_SHORT_LOADS=$mask FENERF_B200_LIB=fenerf_b200/lib_dbg.so timeout 120 python tools/ab_field.py $m 10 2>/dev/null | sed 's/.*median/median/'
done
