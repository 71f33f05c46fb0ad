#!/bin/bash
set -x
mkdir -p gpurun_out
# + PDL on the GEMM kernels (side build 3)
export AQLM_B200_LIB=$PWD/aqlm_b200/csrc/libaqlm_b200.so.new3
timeout 900 python tools/probe_gemm.py --shapes 4096x14336,4096x4096,14336x4096 --batches 256 --settings ";PDL=0;GEMM_DEBUG=12" > gpurun_out/probe_gemm_h1.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --no-check --shapes 64x14336,1024x14336 --batches 256 --settings ";PDL=0" > gpurun_out/probe_gemm_h2.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --op matmat_dequant_transposed --shapes 4096x14336,14336x4096 --batches 256 --settings ";PDL=0" > gpurun_out/probe_gemm_h3.jsonl 2>&1
cat gpurun_out/probe_gemm_h*.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu_h.log 2>&1; tail -3 gpurun_out/pytest_gpu_h.log
# + all-warp epilogue (side build 4)
export AQLM_B200_LIB=$PWD/aqlm_b200/csrc/libaqlm_b200.so.new4
timeout 900 python tools/probe_gemm.py --shapes 4096x14336,4096x4096,14336x4096 --batches 256,64 --settings ";PDL=0" > gpurun_out/probe_gemm_h4.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --no-check --shapes 64x14336,1024x14336 --batches 256 --settings ";GEMM_DEBUG=16" > gpurun_out/probe_gemm_h5.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --op matmat_dequant_transposed --shapes 4096x14336,14336x4096,4096x4096 --batches 256 > gpurun_out/probe_gemm_h6.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --scheme 2x8 --shapes 4096x11008 --batches 256 > gpurun_out/probe_gemm_h7.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --scheme 8x8 --shapes 4096x11008 --batches 256 > gpurun_out/probe_gemm_h8.jsonl 2>&1
cat gpurun_out/probe_gemm_h[4-8].jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu_h4.log 2>&1; tail -3 gpurun_out/pytest_gpu_h4.log
