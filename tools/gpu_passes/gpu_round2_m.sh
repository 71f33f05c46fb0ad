#!/bin/bash
# Re-entry pass M (N=1): full gpu suite at HEAD (LUT cluster second form + 256-bit g16 gathers are the defaults), LUT first/second
# form probe, g16 probe, ncu of the shipped GEMM / LUT cluster kernel, ncu launch list of a short bench step.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_m.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_m.log
tail -6 gpurun_out/pytest_gpu_m.log
timeout 300 python tools/probe_lut2.py > gpurun_out/probe_lut2_m.jsonl 2>&1
cat gpurun_out/probe_lut2_m.jsonl
timeout 200 python - > gpurun_out/probe_g16_m.jsonl 2>&1 <<'PY'
import json, sys, torch
sys.path.insert(0, "tools")
from probe_gemv import make_weights, time_graph
from aqlm_b200.inference_kernels import cuda_kernel
for (fin, fout) in ((4096, 14336), (14336, 4096), (4096, 4096)):
    ws = make_weights(fin, fout, 1, 16, 16, 24, "cuda:0")
    x = torch.randn((1, fin), dtype=torch.float16, device="cuda:0")
    us = time_graph([(lambda w=w: cuda_kernel.matmat(x, w[0], w[1], w[2], None)) for w in ws])
    cb = fout * (fin // 16) * 2
    print(json.dumps(dict(case="1x16 g16 f16, 256-bit gathers", shape=f"{fin}x{fout}", us=round(us, 2), code_GBps=round(cb / us / 1e3, 1))), flush=True)
    del ws
PY
cat gpurun_out/probe_g16_m.jsonl
NCU="ncu --set full --clock-control none --import-source on"
cap() { # name, kernel regex, count, command...
  name=$1; rx=$2; cnt=$3; shift 3
  timeout 240 $NCU -k regex:$rx -c $cnt -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/ncu_$name.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details --csv > gpurun_out/ncu_${name}_details.csv 2>/dev/null
}
cap gemm_f16_final gemm_dequant 1 python tools/ncu_targets.py gemm_f16 1
cap lut_2x8_cluster2 gemv_lut 1 python tools/ncu_targets.py lut_2x8 1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_bench_m.csv python bench.py --steps 2 --warmup 1 --skip-secondary --skip-cpu > gpurun_out/bench_under_ncu_m.log 2>&1
tail -3 gpurun_out/ncu_launches_bench_m.csv | cut -c1-300
du -sh gpurun_out; ls -la gpurun_out
