#!/bin/bash
set -x
mkdir -p gpurun_out
# whole-warp MMA/TMA issue loops (side build 2)
export AQLM_B200_LIB=$PWD/aqlm_b200/csrc/libaqlm_b200.so.new2
timeout 900 python tools/probe_gemm.py --shapes 4096x14336,4096x4096,14336x4096 --batches 256 --settings ";GEMM_DEBUG=12;GEMM_DEBUG=8;GEMM_DEBUG=4;GEMM_CLUSTER=1;GEMM_GROUPS=4;GEMM_A_STAGES=3;GEMM_ATMEM=0" > gpurun_out/probe_gemm_g1.jsonl 2>&1
# fixed cost: 1 / 4 / 16 k-blocks, with and without the epilogue stores
timeout 900 python tools/probe_gemm.py --no-check --shapes 64x14336,256x14336,1024x14336 --batches 256,64 --settings ";GEMM_DEBUG=16;GEMM_DEBUG=28" > gpurun_out/probe_gemm_g2.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --shapes 4096x14336 --batches 64,16 --settings ";GEMM_ATMEM=0" > gpurun_out/probe_gemm_g4.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --scheme 2x8 --shapes 4096x11008 --batches 256 --settings ";GEMM_ATMEM=0" > gpurun_out/probe_gemm_g5.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --scheme 8x8 --shapes 4096x11008 --batches 256 --settings ";GEMM_ATMEM=1" > gpurun_out/probe_gemm_g6.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --op matmat_dequant_transposed --shapes 4096x14336,4096x4096,14336x4096 --batches 256 > gpurun_out/probe_gemm_g7.jsonl 2>&1
cat gpurun_out/probe_gemm_g*.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu_g.log 2>&1; tail -3 gpurun_out/pytest_gpu_g.log
# + PDL on the GEMM kernels (side build 3)
export AQLM_B200_LIB=$PWD/aqlm_b200/csrc/libaqlm_b200.so.new3
timeout 900 python tools/probe_gemm.py --shapes 4096x14336,4096x4096,14336x4096 --batches 256 --settings ";GEMM_DEBUG=12;PDL=0" > gpurun_out/probe_gemm_g8.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --op matmat_dequant_transposed --shapes 4096x14336,14336x4096 --batches 256 > gpurun_out/probe_gemm_g9.jsonl 2>&1
cat gpurun_out/probe_gemm_g8.jsonl gpurun_out/probe_gemm_g9.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu_g3.log 2>&1; tail -3 gpurun_out/pytest_gpu_g3.log
