set -x
CMD="python bench.py --pair-cap-log2 29 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_pairs_changed|k_accumulate|k_expand_bin" -s 18 -c 3 -o gpurun_out/prof_r1 -f $CMD > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
grep bench gpurun_out/plain.log | cut -c1-300
