# N-GPU evidence: the torchrun parity test (sharded run + device exchange vs the C oracle) and the whole-day bench
N=$1
set -x
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo rc=$?
grep -E "Error|error|parity" gpurun_out/r2_bench_${N}gpu.err | tail -5 | cut -c1-300
tail -5 gpurun_out/r2_bench_${N}gpu.err | cut -c1-600
cat gpurun_out/r2_bench_${N}gpu.json
