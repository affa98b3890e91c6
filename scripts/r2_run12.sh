set -x
python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_edge.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -12
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-finalize"
run() { # name, chunk, env...
  name=$1; ch=$2; shift; shift
  env "$@" $B --chunk $ch > gpurun_out/r12_$name.json 2> gpurun_out/r12_$name.err; echo rc=$?; tail -1 gpurun_out/r12_$name.err | cut -c1-300
  python -c "
import json;d=json.load(open('gpurun_out/r12_$name.json'));print('$name', '%.3e'%d['value'], round(d['ms_per_step'],2), d['roofline']['kernels_ms'], d['roofline'].get('random_update_ceiling'))"
}
run c240 240 A=1
run c160 160 A=1
run c120 120 A=1
