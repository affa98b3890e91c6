timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_config3.json 2> gpurun_out/bench_config3.err; echo rc=$?
timeout 300 python bench.py --workload config2 --steps 6 --warmup 3 > gpurun_out/bench_config2.json 2> gpurun_out/bench_config2.err; echo rc=$?
timeout 300 python bench.py --steps 45 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_long.json 2> gpurun_out/bench_long.err; echo rc=$?; tail -2 gpurun_out/bench_long.err | cut -c1-300
