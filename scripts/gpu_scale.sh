N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 12 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo rc=$?; grep -E "Error|error" gpurun_out/bench_${N}gpu.err | tail -3 | cut -c1-300; python -c "
import json;d=json.load(open('gpurun_out/bench_${N}gpu.json'));print(d['n_gpus'],d['value'],d['e2e']['value'],d['final_reduce_ms'])"
