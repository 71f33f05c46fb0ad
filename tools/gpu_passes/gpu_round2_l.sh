#!/bin/bash
# LUT cluster kernel second form + 256-bit gathers for g=16: parity subset, probes
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lut or kx8 or golden or schemes or g16 or bf16 or graph" > gpurun_out/pytest_l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_l.log
tail -5 gpurun_out/pytest_l.log
timeout 600 python tools/probe_lut2.py > gpurun_out/probe_lut2_l.jsonl 2>&1
cat gpurun_out/probe_lut2_l.jsonl
timeout 300 python - > gpurun_out/probe_g16_l.jsonl 2>&1 <<'PY'
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
cat gpurun_out/probe_g16_l.jsonl
