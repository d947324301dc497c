#!/bin/bash
# ncu evidence for one hybrid step at the headline config (1 GPU).  Run under gpurun:
#   gpurun --timeout 1500 -- bash tools/profile_step.sh
# Outputs land in gpurun_out/ (scratch); summarise into profiles/ with profiles/summarize.py / launch_summary.py.
set -u
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c1 --sparse-c4-docs 0 --no-side-configs --parity-queries 0 --synth device --sparse-queries 8"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_step_r2.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_step_r2.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_step_r2.csv $CMD > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_filter -s 12 -c 2 -o gpurun_out/prof_tc_filter_r02 -f $CMD > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bm25_tile -c 1 -o gpurun_out/prof_bm25_tile_r02 -f $CMD > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:tc_select|rescore|tc_finalize|bm25_candidates|fuse_topk" -s 7 -c 5 -o gpurun_out/prof_small_r02 -f $CMD > gpurun_out/ncu_d.log 2>&1
for f in a b c d; do tail -n 2 gpurun_out/ncu_$f.log; done
ls -la gpurun_out/*r02*.ncu-rep
