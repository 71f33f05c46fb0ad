#!/bin/bash
set -x
mkdir -p gpurun_out
S="GEMM_TILE_M=128,GEMM_KSPLIT=4;GEMM_TILE_M=128,GEMM_KSPLIT=3;GEMM_TILE_M=128,GEMM_KSPLIT=5;GEMM_TILE_M=64,GEMM_KSPLIT=2;GEMM_TILE_M=56,GEMM_KSPLIT=2;GEMM_TILE_M=88,GEMM_KSPLIT=3;GEMM_TILE_M=44,GEMM_KSPLIT=1;GEMM_TILE_M=32,GEMM_KSPLIT=1;GEMM_TILE_M=28,GEMM_KSPLIT=1;GEMM_TILE_M=112,GEMM_KSPLIT=4;"
timeout 900 python tools/probe_gemm.py --no-check --shapes 14336x4096,4096x4096 --batches 256 --settings "$S" > gpurun_out/probe_gemm_j1.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --no-check --shapes 14336x4096,4096x4096 --batches 64 --settings "GEMM_TILE_M=128,GEMM_KSPLIT=4;GEMM_TILE_M=64,GEMM_KSPLIT=2;GEMM_TILE_M=32,GEMM_KSPLIT=1;GEMM_TILE_M=128,GEMM_KSPLIT=2;" > gpurun_out/probe_gemm_j2.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --no-check --op matmat_dequant_transposed --shapes 4096x4096,4096x14336 --batches 256 --settings "GEMM_KSPLIT=2;GEMM_KSPLIT=3;GEMM_KSPLIT=4;GEMM_KSPLIT=5;GEMM_KSPLIT=6;" > gpurun_out/probe_gemm_j3.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --no-check --op matmat_dequant_transposed --shapes 14336x4096 --batches 256 --settings "GEMM_KSPLIT=1;GEMM_KSPLIT=2;" > gpurun_out/probe_gemm_j4.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --no-check --shapes 4096x14336 --batches 256,128 --settings ";GEMM_TILE_M=97;GEMM_TILE_M=104;GEMM_TILE_M=112;GEMM_A_STAGES=4;GEMM_STAGES=2" > gpurun_out/probe_gemm_j5.jsonl 2>&1
# (transposed tweaks are in the main build now)
timeout 900 python tools/probe_gemm.py --op matmat_dequant_transposed --shapes 4096x4096,4096x14336,14336x4096 --batches 256 > gpurun_out/probe_gemm_j6.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --dtype bf16 --op matmat_dequant_transposed --shapes 4096x4096 --batches 256 > gpurun_out/probe_gemm_j7.jsonl 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "transposed or backward" 2>&1 | tail -3

cat gpurun_out/probe_gemm_j*.jsonl
