N=$1
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo rc=$?
grep -E "Error|error|parity" gpurun_out/r2_bench_${N}gpu.err | tail -3 | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_${N}gpu.json"))
print("N", d["n_gpus"], "value %.3e e2e %.3e ms/day %.1f parity %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["mgpu_parity"]))
print(d["e2e"]); print(d["day"])
PY
