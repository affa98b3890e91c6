
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --workload config5 --steps 4 --warmup 3 --block 32 --no-e2e > gpurun_out/bench_config5.json 2> gpurun_out/bench_config5.err; echo rc=$?; grep -E "bench|Error" gpurun_out/bench_config5.err | tail -3 | cut -c1-400
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -k "twofloor" 2>&1 | tail -6
