#!/bin/bash
# ncu captures behind profiles/: run on the GPU box (gpurun), then `python profiles/summarize.py <tag>` here.
#   gpurun --timeout 1500 -- 'bash tools/capture_profiles.sh r2f'
TAG=${1:-r2f}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 3 --warmup 3 --kernels-only --no-graphs"
# launch list of one bench invocation (per-kernel time shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/launches_$TAG.log 2>&1
# one full capture per hot kernel
for K in fac_bwd_march fac_fwd_march dcn_fwd_box dcn_bwd_box dcn_gin_collect; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/prof_${K}_$TAG $BENCH > $OUT/prof_${K}_$TAG.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kpn_fused -s 2 -c 1 -f -o $OUT/prof_kpn_fused_$TAG python tools/run_kpn_once.py > $OUT/prof_kpn_fused_$TAG.log 2>&1
# event encoders (BASELINE configs[2])
for K in events_voxel_kernel events_stack_kernel; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o $OUT/prof_${K}_$TAG python tools/run_events_once.py > $OUT/prof_${K}_$TAG.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_events_$TAG.csv python tools/run_events_once.py > /dev/null 2>&1
ls -la $OUT/*_$TAG.ncu-rep
