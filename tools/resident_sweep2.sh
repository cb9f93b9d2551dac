#!/bin/bash
# second pass of tools/resident_sweep.sh: the product rule (ctas_per_sm=0 -> launch_step's choice) against neighbouring bounds
TAG=${1:-r2_v28}
O=gpurun_out
F="--steps 20 --warmup 3 --no-cpu-baseline --no-synth --no-collector --no-e2e --no-per-step --no-packed"
run() {  # config pair ctas
  QG_ENGINE_LIB=$PWD/qiskit_gym_b200/libqg_engine_knobs.so QG_PAIR=$2 QG_REPLAY_CTAS=$3 timeout 200 python bench.py --config $1 $F 2>/dev/null | python -c "
import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 pair=$2 ctas_per_sm=$3 value %.3e frac %.3f ring %d'%(b['value'], b['roofline']['frac'], b['obs_buffers']))"
}
{
for c in C1_perm_grid3 C2_lf8_line C4_pauli10_line; do
  for r in 0 6 7 9 10 12; do run $c 1 $r; done
done
for c in C3_clifford8_full C5_perm27_heavyhex; do
  for r in 0 4 7; do run $c 1 $r; done
done
} | tee $O/${TAG}_resident_sweep2.txt
