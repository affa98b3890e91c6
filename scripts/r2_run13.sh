set -x
python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_edge.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-finalize"
run() { # name, env...
  name=$1; shift
  env "$@" $B > gpurun_out/r13_$name.json 2> gpurun_out/r13_$name.err; echo rc=$?; tail -1 gpurun_out/r13_$name.err | cut -c1-200
  python -c "
import json;d=json.load(open('gpurun_out/r13_$name.json'));print('$name', '%.3e'%d['value'], round(d['ms_per_step'],2), d['roofline']['kernels_ms'], d.get('binned_fraction_of_agent_steps'))"
}
run defer A=1
run nodefer CITAM_DEFER_ACC=0
run defer_ilp4 CITAM_ACC_ILP=4
run defer_ilp1 CITAM_ACC_ILP=1
