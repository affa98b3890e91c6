# N-GPU evidence: the torchrun parity test (sharded run + device exchange vs the C oracle), the whole-day bench
# (strong scaling, exchange inside the timed day) and a slice of the config-4 replica sweep
N=$1
REPL=${2:-128}
set -x
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo rc=$?
grep -E "Error|error|parity" gpurun_out/r2_bench_${N}gpu.err | tail -5 | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_${N}gpu.json"))
print("N", d["n_gpus"], "value %.3e e2e %.3e ms/day %.1f parity %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["mgpu_parity"]))
print(json.dumps(d["day"], indent=1))
print(d["finalize"])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 scripts/sweep_demo.py $REPL 6 > gpurun_out/r2_sweep_${N}gpu.json 2> gpurun_out/r2_sweep_${N}gpu.err; echo rc=$?
cat gpurun_out/r2_sweep_${N}gpu.json; tail -3 gpurun_out/r2_sweep_${N}gpu.err | cut -c1-300
