#!/bin/bash
# One gpurun call: GPU parity tests, smoke, default bench, reference arm, ncu launch list + full capture of the search kernel.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh TAG'
TAG=${1:-r1_vX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/${TAG}_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.txt 2>&1
timeout 600 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_n1.json 2> $O/${TAG}_bench_reference_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-collector --synth-searches 1 > $O/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_search_fused -c 1 -f -o $O/${TAG}_search_fused \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-per-step --no-packed --no-collector --synth-searches 1 > $O/${TAG}_ncu_search.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 3 -c 1 -f -o $O/${TAG}_step \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-per-step --no-packed --no-collector --no-synth > $O/${TAG}_ncu_step.log 2>&1
tail -3 $O/${TAG}_tests.txt; cat $O/${TAG}_smoke.txt | tail -2; head -c 600 $O/${TAG}_bench_n1.json
