#!/bin/bash
# Replay launches (65 536 envs x 128 env-steps, one observation slab per step: every byte of a launch reaches DRAM): CTAs resident per SM
# (QG_REPLAY_CTAS, 0 = whatever fits) x single-warp tiles (pair=0: 2 tiles per CTA) / warp pairs (pair=1: 1 tile per CTA).  Tools build.
TAG=${1:-r2_v27}
O=gpurun_out
F="--steps 20 --warmup 3 --no-cpu-baseline --no-synth --no-collector --no-e2e --no-per-step --no-packed --obs-buffers 128"
run() {  # config pair ctas
  QG_ENGINE_LIB=$PWD/qiskit_gym_b200/libqg_engine_knobs.so QG_PAIR=$2 QG_REPLAY_CTAS=$3 timeout 200 python bench.py --config $1 $F 2>/dev/null | python -c "
import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 pair=$2 ctas_per_sm=$3 value %.3e frac %.3f'%(b['value'], b['roofline']['frac']))"
}
{
for c in C3_clifford8_full C5_perm27_heavyhex; do
  for r in 0 2 3 4 6; do run $c 0 $r; done
  for r in 0 3 4 5 6 8 10; do run $c 1 $r; done
done
for c in C1_perm_grid3 C2_lf8_line C4_pauli10_line; do
  for r in 0 3 5; do run $c 0 $r; done
  for r in 4 8; do run $c 1 $r; done
done
} | tee $O/${TAG}_resident_sweep.txt
