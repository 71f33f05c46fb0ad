#!/bin/bash
# Last pass O (N=1): the shipped defaults after the LUT cluster form selection became automatic -- full gpu suite, smoke, LUT probe.
set -x
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_o.log
tail -4 gpurun_out/pytest_gpu_o.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_o.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_o.log; tail -2 gpurun_out/smoke_o.log
timeout 90 python tools/probe_lut2.py > gpurun_out/probe_lut2_o.jsonl 2>&1
grep -c shipped gpurun_out/probe_lut2_o.jsonl
