#!/bin/bash
# Final validation pass N (N=1): full gpu suite + smoke at the shipped defaults (LUT cluster kernel: first form), then the
# repaired second form behind its switch (parity subset + probe), ncu launch list of a bench step, both bench arms,
# ncu of the LUT cluster kernel and DRAM traffic of the 70B shard-size launches.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_n.txt
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_n.log
tail -4 gpurun_out/pytest_gpu_n.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_n.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_n.log; tail -3 gpurun_out/smoke_n.log
# second form of the LUT cluster kernel (opt-in): parity subset, then first/second form timing
AQLM_B200_LUT_CLUSTER=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "lut or kx8 or golden or schemes or bf16 or graph or flat" > gpurun_out/pytest_lut2_n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_lut2_n.log
tail -4 gpurun_out/pytest_lut2_n.log
timeout 200 python tools/probe_lut2.py > gpurun_out/probe_lut2_n.jsonl 2>&1
cat gpurun_out/probe_lut2_n.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemv -c 300 --csv --log-file gpurun_out/ncu_launches_bench_n.csv python bench.py --steps 2 --warmup 1 --skip-secondary --skip-cpu > gpurun_out/bench_under_ncu_n.log 2>&1
tail -2 gpurun_out/ncu_launches_bench_n.csv | cut -c1-400
( time timeout 540 python bench.py > gpurun_out/bench_n1_n.json 2> gpurun_out/bench_n1_n.err ) 2>&1 | tail -3; echo "bench rc=$?"
tail -3 gpurun_out/bench_n1_n.err
cut -c1-1200 gpurun_out/bench_n1_n.json
( time timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n.json 2> gpurun_out/bench_ref_n.err ) 2>&1 | tail -3
cut -c1-600 gpurun_out/bench_ref_n.json
NCU="ncu --set full --clock-control none --import-source on"
cap() { # name, kernel regex, count, command...
  name=$1; rx=$2; cnt=$3; shift 3
  timeout 200 $NCU -k regex:$rx -c $cnt -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/ncu_$name.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details --csv > gpurun_out/ncu_${name}_details.csv 2>/dev/null
}
cap lut_2x8_cluster gemv_lut 1 python tools/ncu_targets.py lut_2x8 1
cap shards_70b_n8 gemv_1x16 4 python tools/ncu_targets.py shards_70b_n8 1
cap shards_70b_n2 gemv_1x16 4 python tools/ncu_targets.py shards_70b_n2 1
cap shards_70b_n4 gemv_1x16 4 python tools/ncu_targets.py shards_70b_n4 1
du -sh gpurun_out; ls gpurun_out
