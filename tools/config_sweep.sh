#!/bin/bash
# Bench line per BASELINE.json config (replay + per-step + packed + e2e legs) -> gpurun_out/<TAG>_bench_<config>[_t<tile>].json
# usage: bash tools/config_sweep.sh TAG [tile_envs ...]     (tile_envs 0 = automatic)
TAG=${1:-r2_vX}; shift
TILES=${@:-0}
for t in $TILES; do
for c in C1_perm_grid3 C2_lf8_line C3_clifford8_full C4_pauli10_line C5_perm27_heavyhex; do
  timeout 300 python bench.py --config $c --tile-envs $t --steps 20 --warmup 3 --no-cpu-baseline --no-synth --no-collector > gpurun_out/${TAG}_bench_${c}_t${t}.json 2> gpurun_out/${TAG}_bench_${c}_t${t}.err
done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_C*.json")):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_")[1][:-5].ljust(28), "value %.3e"%b["value"], "frac %.3f"%b["roofline"]["frac"], "per_step %.3e %.3f"%(b["per_step_launch"]["value"], b["per_step_launch"]["roofline_frac"]) if b.get("per_step_launch") else "", "e2e %.3e"%b["e2e"]["value"] if b.get("e2e") else "", "packed %.3e"%b["packed_obs"]["value"] if b.get("packed_obs") else "")
    except Exception as e:
        print(f, "ERR", e)
PY
