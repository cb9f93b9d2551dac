#!/bin/bash
# ncu --set full of one replay launch (nsteps = 128) and two single-step launches of k_step for every BASELINE.json config, plus the
# pipe-utilisation metrics BASELINE.json's north_star names (integer ALU / LOP3-class issue, LSU).  One process per config: the bench's
# first k_step launch is the eager replay episode, launches 1.. are the eager single-step launches of the per-step graph capture.
# usage: gpurun --timeout 1500 -- 'bash tools/ncu_configs.sh TAG [configs...]'      -> gpurun_out/<TAG>_<config>.ncu-rep + _raw.csv
TAG=${1:-r2_vX}; shift
CONFIGS=${@:-C1_perm_grid3 C2_lf8_line C3_clifford8_full C4_pauli10_line C5_perm27_heavyhex}
O=gpurun_out
mkdir -p $O
EXTRA=smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_uniform.sum,smsp__inst_executed_pipe_adu.sum,smsp__inst_executed_pipe_cbu.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_generic_atom_dot_alu.sum,smsp__sass_inst_executed_op_shared_ld.sum,smsp__sass_inst_executed_op_shared_st.sum,smsp__sass_inst_executed_op_global_st.sum,smsp__sass_inst_executed_op_global_ld.sum
for c in $CONFIGS; do
  timeout 420 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:k_step -s 0 -c 3 -f -o $O/${TAG}_${c} \
    python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-packed --no-collector --no-synth > $O/${TAG}_${c}_ncu.log 2>&1
  python tools/ncu_summarize_box.py $O/${TAG}_${c}.ncu-rep $O/${TAG}_${c}_ncu.txt 2> $O/${TAG}_${c}_sum.err
  rm -f $O/${TAG}_${c}.ncu-rep          # (tens of MB each: gpurun_out is capped at 64 MiB; the summary holds what profiles/ keeps)
  tail -1 $O/${TAG}_${c}_ncu.log | head -c 300; echo
done
ls -la $O | grep ${TAG} | head -30
