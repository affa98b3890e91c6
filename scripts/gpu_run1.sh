set -x
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python bench.py --workload config2 --steps 6 --warmup 3 > gpurun_out/bench_config2.json 2> gpurun_out/bench_config2.err; echo rc=$?; grep bench gpurun_out/bench_config2.err | cut -c1-400; cat gpurun_out/bench_config2.json
timeout 600 python bench.py --steps 12 --warmup 3 > gpurun_out/bench_config3.json 2> gpurun_out/bench_config3.err; echo rc=$?; tail -4 gpurun_out/bench_config3.err | cut -c1-400; cat gpurun_out/bench_config3.json
