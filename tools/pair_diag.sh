#!/bin/bash
# Replay A/B of the warp-pair variants (tools builds) over two observation-ring sizes: ring 0 = the bench's automatic ring (5 slabs for C3: a tile
# rewrites a slab every 5 steps), ring 128 = one slab per step (every byte of a launch has to reach DRAM).
# knobs = product kernels (named barriers, immediate numbers; make EXTRA=-DQG_TOOLS_KNOBS), knobs_mb = mbarriers + relaxed arrive (EXTRA="-DQG_TOOLS_KNOBS
# -DQG_PAIR_MBARRIER"), knobs_dyn = the first pair kernel: barrier number in a register, which made ptxas reserve 16 barriers per CTA (4 CTAs per SM) — that
# variant was removed from the tree after this experiment (profiles/r2_v26_pair_ab.txt); QG_REPLAY_CTAS=4 with the product kernels is its equivalent.
TAG=${1:-r2_v26}
O=gpurun_out
F="--steps 20 --warmup 3 --no-cpu-baseline --no-synth --no-collector --no-e2e --no-per-step --no-packed"
run() {  # config lib pair ring
  QG_ENGINE_LIB=$PWD/qiskit_gym_b200/libqg_engine_$2.so QG_PAIR=$3 timeout 200 python bench.py --config $1 --obs-buffers $4 $F 2>/dev/null | python -c "
import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2 pair=$3 ring=$4 value %.3e frac %.3f'%(b['value'], b['roofline']['frac']))"
}
{
for c in C3_clifford8_full C5_perm27_heavyhex; do
for ring in 0 128; do
  run $c knobs 0 $ring; run $c knobs 1 $ring; run $c knobs_dyn 1 $ring; run $c knobs_mb 1 $ring
done; done
for c in C1_perm_grid3 C2_lf8_line C4_pauli10_line; do run $c knobs 0 0; run $c knobs 1 0; run $c knobs_mb 1 0; done
} | tee $O/${TAG}_pair_ab.txt
