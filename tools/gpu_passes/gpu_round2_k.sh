#!/bin/bash
# Re-entry validation pass (N=1): full gpu test suite, smoke, both bench arms, launch list of the bench command
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_k.log
tail -5 gpurun_out/pytest_gpu_k.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_k.log 2>&1; tail -3 gpurun_out/smoke_k.log
( time timeout 900 python bench.py > gpurun_out/bench_n1_k.json 2> gpurun_out/bench_n1_k.err ) 2>&1 | tail -3; echo "bench rc=$?"
tail -3 gpurun_out/bench_n1_k.err
cut -c1-1500 gpurun_out/bench_n1_k.json
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_k.json 2> gpurun_out/bench_ref_k.err ) 2>&1 | tail -3
cut -c1-800 gpurun_out/bench_ref_k.json
