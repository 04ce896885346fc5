#!/bin/bash
# Ablation timings of the tcgen05 point-network launch (variant builds, WRONG results by design).  Build the variants with
#   for v in 1 2 3 4 6 7 8 16 32 40; do FENERF_NVCC_DEFINES=-DFENERF_ABLATE=$v FENERF_B200_LIB=fenerf_b200/lib_ab$v.so python -m fenerf_b200.build; done
# mask bits: 1 short weight loads, 2 FiLM epilogue hand-offs only, 4 no tcgen05.mma, 8 no sin, 16 no activation stores, 32 no tcgen05.ld
#   bash tools/ablate_fast3.sh [A|B] [variant names ...]
m=${1:-A}; shift
libs=${@:-"fenerf_b200 ab1 ab2 ab3 ab4 ab6 ab7 ab8 ab16 ab32 ab40"}
for v in $libs; do
  printf "%-12s " $v
  FENERF_B200_LIB=fenerf_b200/lib$v.so timeout 120 python tools/ab_field.py $m 10 2>/dev/null | sed 's/.*median/median/'
done
