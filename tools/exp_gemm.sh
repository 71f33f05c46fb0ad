#!/bin/bash
for gm in 0 1 3; do
  for st in 2 3; do
    echo "gather_mode=$gm stages=$st"
    AQLM_B200_GEMM_GATHER_MODE=$gm AQLM_B200_GEMM_STAGES=$st timeout 120 python tools/probe_gemv.py --op matmat_dequant --schemes 1x16 --batches 64,256 --shapes 4096x14336,4096x4096 2>&1 | cut -c1-140
  done
done
