#!/bin/bash
# First GPU pass of round 2: parity tests, bench line, ncu evidence of the shipped kernels (run under gpurun).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
# GEMM experiments first (short): A-in-TMEM / ragged tile heights vs the round-1 configuration
timeout 600 python tools/probe_gemm.py --shapes 4096x14336,4096x4096 --batches 256 --settings "GEMM_ATMEM=0,GEMM_TILE_M=128;GEMM_ATMEM=0;GEMM_ATMEM=1,GEMM_TILE_M=128;GEMM_ATMEM=1;GEMM_ATMEM=1,GEMM_TILE_M=97,GEMM_STAGES=2;GEMM_ATMEM=1,GEMM_GATHER_MODE=0" > gpurun_out/probe_gemm_a.jsonl 2>&1
cat gpurun_out/probe_gemm_a.jsonl
export AQLM_B200_GEMM_ATMEM=0
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gemm_dequant -c 2 -o gpurun_out/prof_gemm_f16 python tools/ncu_targets.py gemm_f16 2 > gpurun_out/ncu_gemm_f16.log 2>&1
timeout 300 $NCU -k regex:gemm_dequant -c 1 -o gpurun_out/prof_gemm_bf16 python tools/ncu_targets.py gemm_bf16 1 > gpurun_out/ncu_gemm_bf16.log 2>&1
timeout 300 $NCU -c 10 -o gpurun_out/prof_gather_microbench tools/bin/gather_microbench ncu > gpurun_out/ncu_microbench.log 2>&1
timeout 300 $NCU -k regex:gemv_lut -c 2 -o gpurun_out/prof_lut_2x8 python tools/ncu_targets.py lut_2x8 2 > gpurun_out/ncu_lut.log 2>&1
timeout 200 tools/bin/gather_microbench > gpurun_out/gather_microbench.jsonl 2>&1
ls -la gpurun_out | head -40
