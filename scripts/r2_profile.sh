# ncu evidence for the current build: per-launch device times of one whole day, and a full capture of
# every kernel of one mid-day chunk.  Run under gpurun; summaries go to profiles/ (scripts/ncu_summary.py).
set -x
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-finalize"
$CMD > gpurun_out/r2_prof_plain.json 2> gpurun_out/r2_prof_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 2130 -c 720 --csv --log-file gpurun_out/r2_launches_day.csv $CMD > gpurun_out/r2_ncu1.log 2>&1
echo launches rc=$?
$CMD > gpurun_out/r2_prof_plain2.json 2> gpurun_out/r2_prof_plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:'k_classify|k_expand_dyn|k_scan|k_scatter|k_select|k_pairs|k_accumulate' -s 300 -c 11 -o gpurun_out/r2_prof_chunk -f $CMD > gpurun_out/r2_ncu2.log 2>&1
echo full rc=$?
