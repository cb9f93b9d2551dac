#!/bin/bash
# Single-step launches (qg_step in a CUDA graph, 65 536 envs): CTAs resident per SM (QG_STEP_CTAS, tools build; 0 = whatever fits).
TAG=${1:-r2_v55}
O=gpurun_out
F="--steps 10 --warmup 3 --no-cpu-baseline --no-synth --no-collector --no-e2e --no-packed"
run() {  # config ctas
  QG_ENGINE_LIB=$PWD/qiskit_gym_b200/libqg_engine_knobs.so QG_STEP_CTAS=$2 timeout 200 python bench.py --config $1 $F 2>/dev/null | python -c "
import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=b['per_step_launch']; print('$1 step_ctas_per_sm=$2 per-step %.3e frac %.3f launch %.2f us'%(p['value'], p['roofline_frac'], p['avg_launch_us']))"
}
{
for c in C3_clifford8_full C5_perm27_heavyhex C4_pauli10_line C1_perm_grid3; do
  for r in 0 3 4 5 7; do run $c $r; done
done
} | tee $O/${TAG}_step_resident_sweep.txt
