set -x
timeout 300 python bench.py --workload config2 --steps 6 --warmup 3 > gpurun_out/bench_config2.json 2> gpurun_out/bench_config2.err; echo rc=$?; grep bench gpurun_out/bench_config2.err | cut -c1-300; python -c "
import json;d=json.load(open('gpurun_out/bench_config2.json'));print(d['value'],d['e2e'],d['roofline']['kernels_ms'])"
timeout 600 python bench.py --steps 12 --warmup 3 > gpurun_out/bench_config3.json 2> gpurun_out/bench_config3.err; echo rc=$?; grep bench gpurun_out/bench_config3.err | cut -c1-300; python -c "
import json;d=json.load(open('gpurun_out/bench_config3.json'));print(d['value'],d['e2e'],d['roofline']['kernels_ms'])"
