set -x
python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_edge.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -6
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-finalize"
run() { # name, chunk
  name=$1; ch=$2; shift; shift
  env "$@" $B --chunk $ch > gpurun_out/r14_$name.json 2> gpurun_out/r14_$name.err; echo rc=$?; tail -1 gpurun_out/r14_$name.err | cut -c1-200
  python -c "
import json;d=json.load(open('gpurun_out/r14_$name.json'));print('$name', '%.3e'%d['value'], round(d['ms_per_step'],2), d['roofline']['kernels_ms'], d.get('binned_fraction_of_agent_steps'))"
}
run c480 480 A=1
run c360 360 A=1
run c240 240 A=1
run c0 0 A=1
