nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 12 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo rc=$?; grep -E "Error|error" gpurun_out/bench_8gpu.err | tail -5 | cut -c1-300; cut -c1-300 gpurun_out/bench_8gpu.json
