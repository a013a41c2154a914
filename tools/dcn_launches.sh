#!/bin/bash
# DCN fwd/bwd timing at cfg1 + per-kernel durations of one call (ncu launch list)
python tools/time_dcn.py
python tools/check_dcn_det.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/l.csv python tools/run_dcn_once.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/l.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
for r in rows[1:][-8:]: print(r[ki][:60], r[vi])
PY
