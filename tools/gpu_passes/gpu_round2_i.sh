#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python tools/probe_gemm.py --shapes 4096x14336,4096x4096,14336x4096 --batches 256,64 --settings ";PDL=0" > gpurun_out/probe_gemm_i1.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --no-check --shapes 64x14336,1024x14336 --batches 256 --settings ";GEMM_DEBUG=16" > gpurun_out/probe_gemm_i2.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --op matmat_dequant_transposed --shapes 4096x14336,14336x4096,4096x4096 --batches 256 > gpurun_out/probe_gemm_i3.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --scheme 2x8 --shapes 4096x11008 --batches 256 > gpurun_out/probe_gemm_i4.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --scheme 8x8 --shapes 4096x11008 --batches 256 > gpurun_out/probe_gemm_i5.jsonl 2>&1
timeout 900 python tools/probe_gemm.py --dtype bf16 --shapes 4096x14336 --batches 256 > gpurun_out/probe_gemm_i6.jsonl 2>&1
cat gpurun_out/probe_gemm_i*.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_i.log
tail -5 gpurun_out/pytest_gpu_i.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_i.log 2>&1; tail -3 gpurun_out/smoke_i.log
# LUT cluster kernel, 16-row warp batches on 768 threads (side build)
export AQLM_B200_LIB=$PWD/aqlm_b200/csrc/libaqlm_b200.so.rb16
timeout 300 python tools/probe_lut.py 2>&1 | grep -E '"full"' > gpurun_out/probe_lut_i_rb32.jsonl
AQLM_B200_LUT_RB16=1 timeout 300 python tools/probe_lut.py 2>&1 | grep -E '"full"' > gpurun_out/probe_lut_i_rb16.jsonl
cat gpurun_out/probe_lut_i_rb32.jsonl gpurun_out/probe_lut_i_rb16.jsonl
AQLM_B200_LUT_RB16=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lut or kx8 or golden or schemes" 2>&1 | tail -3
