#!/bin/bash
# One gpurun call: GPU parity tests, smoke, default bench, reference arm, 1-GPU reference of the multi-GPU differential test.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh TAG [quick]'
TAG=${1:-r2_vX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/${TAG}_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -25 > $O/${TAG}_tests.txt
tail -4 $O/${TAG}_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.txt 2>&1
tail -2 $O/${TAG}_smoke.txt
if [ "$2" != "quick" ]; then
timeout 600 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
head -c 900 $O/${TAG}_bench_n1.json; echo; tail -3 $O/${TAG}_bench_n1.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_n1.json 2> $O/${TAG}_bench_reference_n1.err
head -c 400 $O/${TAG}_bench_reference_n1.json; echo
timeout 300 python tools/multi_gpu_parity.py --out $O/${TAG}_parity > $O/${TAG}_parity_n1.log 2>&1
tail -2 $O/${TAG}_parity_n1.log
fi
