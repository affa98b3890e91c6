nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 6 --warmup 3 --no-e2e > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo rc=$?; grep -E "bench|Error|error" gpurun_out/bench_2gpu.err | tail -5 | cut -c1-300; cut -c1-400 gpurun_out/bench_2gpu.json
