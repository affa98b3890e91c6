set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_config3.json 2> gpurun_out/bench_config3.err; echo rc=$?
timeout 300 python bench.py --workload config2 --steps 6 --warmup 3 > gpurun_out/bench_config2.json 2> gpurun_out/bench_config2.err; echo rc=$?
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?
CMD="python bench.py --pair-cap-log2 29 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_pairs_changed|k_accumulate|k_expand_bin|k_scatter" -s 24 -c 4 -o gpurun_out/prof_r1 -f $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
