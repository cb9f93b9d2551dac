#!/bin/bash
# ncu launch list (device time per launch) of the packed tensor-core collector, eager: which kernels a decision is made of
TAG=${1:-r2_vX}
O=gpurun_out
timeout 300 python tools/collector_probe.py 65536 32 > $O/${TAG}_collector_probe.json 2> $O/${TAG}_collector_probe.err; cat $O/${TAG}_collector_probe.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file $O/${TAG}_collector_launches.csv \
  python tools/collector_probe.py 65536 8 eager > $O/${TAG}_collector_ncu.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("$O/${TAG}_collector_launches.csv")) if len(r)>10]
h=rows[0]; kn=h.index("Kernel Name"); mv=h.index("Metric Value")
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[kn][:70],[]).append(float(r[mv].replace(",","")))
for k,v in d.items(): print(f"{k:72s} n={len(v):3d} mean_us={sum(v)/len(v)/1e3:8.2f}")
PY
