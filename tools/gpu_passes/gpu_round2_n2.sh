#!/bin/bash
# 2-GPU pass: peer tests (fused GEMV+exchange, two-kernel form, grouped, graph, two devices in one process) and bench --gpus 2
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_n2.txt
timeout 900 python -m pytest tests/test_gpu_peer.py -m gpu -q -x > gpurun_out/pytest_peer.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_peer.log
tail -15 gpurun_out/pytest_peer.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
tail -5 gpurun_out/bench_n2.err
cat gpurun_out/bench_n2.json | cut -c1-3000
AQLM_B200_FUSED_EXCHANGE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --skip-n1 --skip-pairing > gpurun_out/bench_n2_twokernel.json 2> gpurun_out/bench_n2_twokernel.err; echo "bench n2 (two-kernel exchange) rc=$?"
cat gpurun_out/bench_n2_twokernel.json | cut -c1-600
