#!/bin/bash
# ncu launch list + full capture of the tensor-core policy kernels (tools/tc_policy_probe.py)
TAG=${1:-r2_vX}
O=gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_tc -c 24 --csv --log-file $O/${TAG}_tc_launches.csv python tools/tc_policy_probe.py 65536 > $O/${TAG}_tc_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_tc_fused3 -s 2 -c 1 -f -o $O/${TAG}_tc_fused python tools/tc_policy_probe.py 65536 > $O/${TAG}_tc_ncu2.log 2>&1
python tools/ncu_summarize_box.py $O/${TAG}_tc_fused.ncu-rep $O/${TAG}_tc_fused_ncu.txt 2> $O/${TAG}_tc_sum.err
rm -f $O/${TAG}_tc_fused.ncu-rep
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("$O/${TAG}_tc_launches.csv")) if len(r)>10]
h=rows[0]; kn=h.index("Kernel Name"); mv=h.index("Metric Value")
for r in rows[1:9]: print(r[kn][:60], r[mv])
PY
