#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/probe_gemm.py --shapes 4096x14336 --batches 256 --settings "GEMM_ATMEM=1;GEMM_ATMEM=1,GEMM_STAGES=4;GEMM_ATMEM=1,GEMM_STAGES=4,GEMM_CLUSTER=2;GEMM_ATMEM=1,GEMM_CLUSTER=2;GEMM_ATMEM=1,GEMM_STAGES=4,GEMM_TILE_M=97;GEMM_ATMEM=1,GEMM_STAGES=4,GEMM_TILE_M=128" > gpurun_out/probe_gemm_c1.jsonl 2>&1
timeout 600 python tools/probe_gemm.py --scheme 2x8 --shapes 4096x11008 --batches 256 --settings "GEMM_TILE_M=128;GEMM_TILE_M=0;GEMM_TILE_M=128,GEMM_ATMEM=1;GEMM_TILE_M=128,GEMM_V2=1;GEMM_TILE_M=96" > gpurun_out/probe_gemm_c2.jsonl 2>&1
timeout 600 python tools/probe_gemm.py --scheme 8x8 --shapes 4096x11008 --batches 256 --settings "GEMM_TILE_M=128;GEMM_TILE_M=0;GEMM_TILE_M=128,GEMM_ATMEM=1" > gpurun_out/probe_gemm_c3.jsonl 2>&1
cat gpurun_out/probe_gemm_c*.jsonl
timeout 400 python tools/probe_lut.py > gpurun_out/probe_lut_c.jsonl 2>&1
grep -E "full|no fix-up" gpurun_out/probe_lut_c.jsonl
timeout 200 tools/bin/lut16_microbench > gpurun_out/lut16_microbench.jsonl 2>&1
cat gpurun_out/lut16_microbench.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_c.log
tail -8 gpurun_out/pytest_gpu_c.log
