#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lut or kx8 or gemm or golden or schemes" > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_d.log
tail -8 gpurun_out/pytest_gpu_d.log
timeout 400 python tools/probe_lut.py > gpurun_out/probe_lut_d.jsonl 2>&1
grep -E "full|1 CTA" gpurun_out/probe_lut_d.jsonl
AQLM_B200_LUT_CLUSTER=0 timeout 400 python tools/probe_lut.py 2>&1 | grep -E "full" > gpurun_out/probe_lut_d_nocluster.jsonl
cat gpurun_out/probe_lut_d_nocluster.jsonl
timeout 600 python tools/probe_gemm.py --shapes 4096x14336,4096x4096,14336x4096 --batches 256 --settings ";GEMM_CLUSTER=1" > gpurun_out/probe_gemm_d1.jsonl 2>&1
timeout 600 python tools/probe_gemm.py --scheme 2x8 --shapes 4096x11008,4096x4096 --batches 256 --settings ";GEMM_ATMEM=1;GEMM_V2=1;GEMM_TILE_M=128" > gpurun_out/probe_gemm_d2.jsonl 2>&1
timeout 600 python tools/probe_gemm.py --scheme 8x8 --shapes 4096x11008 --batches 256 --settings ";GEMM_ATMEM=1;GEMM_V2=0" > gpurun_out/probe_gemm_d3.jsonl 2>&1
timeout 600 python tools/probe_gemm.py --scheme 1x8 --shapes 4096x11008 --batches 256 --settings ";GEMM_ATMEM=1;GEMM_V2=1" > gpurun_out/probe_gemm_d4.jsonl 2>&1
cat gpurun_out/probe_gemm_d*.jsonl
# decoupled A/X pipelines (side build): A stages in TMEM with their own barriers
export AQLM_B200_LIB=$PWD/aqlm_b200/csrc/libaqlm_b200.so.new
timeout 600 python tools/probe_gemm.py --shapes 4096x14336 --batches 256 --settings ";GEMM_A_STAGES=3;GEMM_A_STAGES=8;GEMM_GROUPS=4;GEMM_GROUPS=4,GEMM_A_STAGES=8;GEMM_DEBUG=8;GEMM_DEBUG=4;GEMM_DEBUG=12;GEMM_CLUSTER=1" > gpurun_out/probe_gemm_d5.jsonl 2>&1
timeout 600 python tools/probe_gemm.py --shapes 4096x4096,14336x4096 --batches 256,64 --settings ";GEMM_GROUPS=4" > gpurun_out/probe_gemm_d6.jsonl 2>&1
cat gpurun_out/probe_gemm_d5.jsonl gpurun_out/probe_gemm_d6.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gemm" > gpurun_out/pytest_gpu_d2.log 2>&1; tail -3 gpurun_out/pytest_gpu_d2.log
