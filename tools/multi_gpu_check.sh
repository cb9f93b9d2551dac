#!/bin/bash
# Multi-GPU evidence in one gpurun --gpus N call: the differential test at 1 GPU (reference) and at every N in the list, then bench.py at the
# largest N.   usage: gpurun --gpus 8 --timeout 1500 -- 'bash tools/multi_gpu_check.sh TAG 2 4 8'
TAG=${1:-r2_vX}
NS=${2:-2}            # GPU counts of the differential test, e.g. "2 4 8"
BN=${3:-$NS}          # GPU counts of the bench runs, e.g. "8"
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1
timeout 300 python tools/multi_gpu_parity.py --out $O/${TAG}_parity > $O/${TAG}_parity_n1.log 2>&1; tail -1 $O/${TAG}_parity_n1.log
LAST=1
for n in $NS; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_parity.py --out $O/${TAG}_parity > $O/${TAG}_parity_n$n.log 2>&1
  echo "parity N=$n rc=$?"; tail -1 $O/${TAG}_parity_n$n.log | head -c 600; echo
  LAST=$n
done
for n in $BN; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_n$n.json 2> $O/${TAG}_bench_n$n.err
echo "bench N=$n rc=$?"
python - <<PY
import json
try:
    b=json.loads(open("$O/${TAG}_bench_n$n.json").read().strip().splitlines()[-1])
    print("N=$n value %.3e e2e %.3e flags_only %.3e int32fmt %.3e synth weak %.0f strong %.0f"%(b["value"], b["e2e"]["value"], b["e2e"]["flags_only_value"] or 0, b["e2e"]["int32_u8_format_value"], b["synth"]["value"], b["synth"]["strong"]["value"] or 0))
except Exception as e:
    print("ERR", e)
PY
done
