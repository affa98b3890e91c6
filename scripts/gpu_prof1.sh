set -x
CMD="python bench.py --workload config2 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_config2.csv $CMD > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pairs -s 6 -c 2 -o gpurun_out/prof_pairs_config2 -f $CMD > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
timeout 600 python bench.py --steps 12 --warmup 3 > gpurun_out/bench_config3.json 2> gpurun_out/bench_config3.err; echo rc=$?; tail -4 gpurun_out/bench_config3.err | cut -c1-400; cat gpurun_out/bench_config3.json
