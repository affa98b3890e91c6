set -x
CMD="python bench.py --workload config3 --agents 400000 --block 128 --pair-cap-log2 27 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pairs -s 3 -c 1 -o gpurun_out/prof_pairs -f $CMD > gpurun_out/ncu_pairs.log 2>&1
tail -2 gpurun_out/ncu_pairs.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 3 -c 1 -o gpurun_out/prof_accum -f $CMD > gpurun_out/ncu_accum.log 2>&1
tail -2 gpurun_out/ncu_accum.log
cat gpurun_out/plain.log | tail -1 | cut -c1-300
