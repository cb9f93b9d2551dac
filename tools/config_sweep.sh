#!/bin/bash
# Bench line per BASELINE.json config (replay + per-step + packed legs only) -> gpurun_out/<TAG>_bench_<config>.json
TAG=${1:-r1_vX}
for c in C1_perm_grid3 C2_lf8_line C4_pauli10_line C5_perm27_heavyhex; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-synth --no-collector > gpurun_out/${TAG}_bench_${c}.json 2> gpurun_out/${TAG}_bench_${c}.err
done
timeout 300 python bench.py --config C3_clifford8_full --envs 1048576 --steps 5 --warmup 3 --no-cpu-baseline --no-synth --no-collector --no-e2e > gpurun_out/${TAG}_bench_C3_1M.json 2> gpurun_out/${TAG}_bench_C3_1M.err
timeout 300 python bench.py --config C3_clifford8_full --add-inverts 1 --steps 10 --warmup 3 --no-cpu-baseline --no-synth --no-collector > gpurun_out/${TAG}_bench_C3_inv.json 2> gpurun_out/${TAG}_bench_C3_inv.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_C*.json")):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_")[1][:-5], "value %.3e"%b["value"], "frac %.3f"%b["roofline"]["frac"], "per_step %.3e %.3f"%(b["per_step_launch"]["value"], b["per_step_launch"]["roofline_frac"]) if b.get("per_step_launch") else "", "e2e %.3e"%b["e2e"]["value"] if b.get("e2e") else "", "packed %.3e"%b["packed_obs"]["value"] if b.get("packed_obs") else "")
    except Exception as e:
        print(f, "ERR", e)
PY
