#!/bin/bash
# ncu evidence for one hybrid step at the headline config (1 GPU).  Run under gpurun:
#   gpurun --timeout 1500 -- bash tools/profile_step.sh
# Outputs land in gpurun_out/ (scratch); summarise into profiles/ with profiles/summarize.py.
set -u
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sparse-queries 8"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_step.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_step.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step.csv $CMD > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_filter -s 12 -c 2 -o gpurun_out/prof_tc_filter_r2 -f $CMD > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bm25_tile -c 1 -o gpurun_out/prof_bm25_tile_r2 -f $CMD > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:tc_select|rescore|tc_finalize|bm25_candidates|fuse_topk" -s 7 -c 5 -o gpurun_out/prof_small_r2 -f $CMD > gpurun_out/ncu_d.log 2>&1
for f in a b c d; do tail -n 2 gpurun_out/ncu_$f.log; done
ls -la gpurun_out/*.ncu-rep
