#!/bin/bash
# Second GPU pass: GEMM follow-up probes, new kernels' tests, bench line, ncu evidence exported to CSV on the box.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 600 python tools/probe_gemm.py --shapes 4096x14336 --batches 256 --settings "GEMM_ATMEM=1;GEMM_ATMEM=1,GEMM_CLUSTER=2;GEMM_ATMEM=1,GEMM_CLUSTER=4;GEMM_ATMEM=0,GEMM_CLUSTER=2;GEMM_ATMEM=1,GEMM_DEBUG=4;GEMM_ATMEM=1,GEMM_DEBUG=8;GEMM_ATMEM=1,GEMM_STAGES=2,GEMM_CLUSTER=2" > gpurun_out/probe_gemm_b1.jsonl 2>&1
timeout 600 python tools/probe_gemm.py --shapes 4096x4096 --batches 256 --settings "GEMM_ATMEM=0;GEMM_ATMEM=1;GEMM_ATMEM=1,GEMM_TILE_M=32,GEMM_KSPLIT=1;GEMM_ATMEM=1,GEMM_TILE_M=28,GEMM_KSPLIT=1;GEMM_ATMEM=1,GEMM_TILE_M=56,GEMM_KSPLIT=2;GEMM_ATMEM=1,GEMM_TILE_M=64,GEMM_KSPLIT=2;GEMM_ATMEM=0,GEMM_TILE_M=32,GEMM_KSPLIT=1;GEMM_ATMEM=1,GEMM_TILE_M=128,GEMM_KSPLIT=2" > gpurun_out/probe_gemm_b2.jsonl 2>&1
timeout 600 python tools/probe_gemm.py --shapes 14336x4096,4096x14336 --batches 64,16 --settings "GEMM_ATMEM=0;GEMM_ATMEM=1" > gpurun_out/probe_gemm_b3.jsonl 2>&1
timeout 600 python tools/probe_gemm.py --scheme 2x8 --shapes 4096x11008 --batches 256 --settings "GEMM_ATMEM=0;GEMM_ATMEM=1;GEMM_ATMEM=0,GEMM_V2=1" > gpurun_out/probe_gemm_b4.jsonl 2>&1
timeout 600 python tools/probe_gemm.py --op matmat_dequant_transposed --shapes 4096x14336,4096x4096,14336x4096 --batches 256 > gpurun_out/probe_gemm_t.jsonl 2>&1
cat gpurun_out/probe_gemm_b*.jsonl gpurun_out/probe_gemm_t.jsonl
timeout 400 python tools/probe_lut.py > gpurun_out/probe_lut.jsonl 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
NCU="ncu --set full --clock-control none --import-source on"
cap() { # name, kernel regex, count, command...
  name=$1; rx=$2; cnt=$3; shift 3
  timeout 300 $NCU -k regex:$rx -c $cnt -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/ncu_$name.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details --csv > gpurun_out/ncu_${name}_details.csv 2>/dev/null
}
cap gemm_f16 gemm_dequant 1 python tools/ncu_targets.py gemm_f16 1
cap gemm_bf16 gemm_dequant 1 python tools/ncu_targets.py gemm_bf16 1
cap gemm_t gemm_dequant_t 1 python tools/ncu_targets.py gemm_t 1
cap lut_2x8 gemv_lut 1 python tools/ncu_targets.py lut_2x8 1
cap layer_1x16 gemv_1x16 4 python tools/ncu_targets.py layer_1x16 1
cap gather_microbench "k_" 5 tools/bin/gather_microbench ncu
timeout 200 tools/bin/gather_microbench > gpurun_out/gather_microbench.jsonl 2>&1
du -sh gpurun_out; ls -la gpurun_out
