#!/bin/bash
set -x
mkdir -p gpurun_out
export AQLM_B200_LIB=$PWD/aqlm_b200/csrc/libaqlm_b200.so.new
# slope/intercept of the pipeline skeleton: same out_features, 32 / 64 / 128 k-blocks, with and without memory traffic
timeout 900 python tools/probe_gemm.py --no-check --shapes 2048x14336,4096x14336,8192x14336 --batches 256 --settings "GEMM_DEBUG=12;GEMM_DEBUG=12,GEMM_CLUSTER=1;GEMM_DEBUG=12,GEMM_CLUSTER=1,GEMM_A_STAGES=3;GEMM_DEBUG=8;GEMM_DEBUG=4;;GEMM_CLUSTER=1" > gpurun_out/probe_gemm_f1.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --no-check --shapes 4096x14336 --batches 128,64 --settings "GEMM_DEBUG=12;GEMM_DEBUG=12,GEMM_CLUSTER=1;" > gpurun_out/probe_gemm_f2.jsonl 2>&1
cat gpurun_out/probe_gemm_f*.jsonl
# whole-warp MMA/TMA issue loops (side build 2)
export AQLM_B200_LIB=$PWD/aqlm_b200/csrc/libaqlm_b200.so.new2
timeout 900 python tools/probe_gemm.py --shapes 4096x14336,4096x4096,14336x4096 --batches 256 --settings ";GEMM_DEBUG=12;GEMM_DEBUG=8;GEMM_DEBUG=4;GEMM_CLUSTER=1;GEMM_GROUPS=4;GEMM_A_STAGES=3;GEMM_ATMEM=0" > gpurun_out/probe_gemm_f3.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --shapes 4096x14336 --batches 64,16 --settings ";GEMM_ATMEM=0" > gpurun_out/probe_gemm_f4.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --scheme 2x8 --shapes 4096x11008 --batches 256 --settings ";GEMM_ATMEM=0" > gpurun_out/probe_gemm_f5.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --scheme 8x8 --shapes 4096x11008 --batches 256 --settings ";GEMM_ATMEM=1" > gpurun_out/probe_gemm_f6.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --op matmat_dequant_transposed --shapes 4096x14336,4096x4096,14336x4096 --batches 256 > gpurun_out/probe_gemm_f7.jsonl 2>&1
cat gpurun_out/probe_gemm_f[3-7].jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu_f.log 2>&1; tail -3 gpurun_out/pytest_gpu_f.log
