#!/bin/bash
set -x
mkdir -p gpurun_out
# decoupled A/X pipelines (side build): A stages in TMEM with their own barriers
export AQLM_B200_LIB=$PWD/aqlm_b200/csrc/libaqlm_b200.so.new
timeout 600 python tools/probe_gemm.py --shapes 4096x14336 --batches 256 --settings ";GEMM_A_STAGES=3;GEMM_A_STAGES=8;GEMM_GROUPS=4;GEMM_GROUPS=4,GEMM_A_STAGES=8;GEMM_DEBUG=8;GEMM_DEBUG=4;GEMM_DEBUG=12;GEMM_CLUSTER=1" > gpurun_out/probe_gemm_e1.jsonl 2>&1
timeout 600 python tools/probe_gemm.py --shapes 4096x4096,14336x4096 --batches 256,64 --settings ";GEMM_GROUPS=4" > gpurun_out/probe_gemm_e2.jsonl 2>&1
timeout 600 python tools/probe_gemm.py --scheme 2x8 --shapes 4096x11008,4096x4096 --batches 256 --settings ";GEMM_GROUPS=4" > gpurun_out/probe_gemm_e3.jsonl 2>&1
cat gpurun_out/probe_gemm_e*.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu_e.log 2>&1; tail -3 gpurun_out/pytest_gpu_e.log
