set -x
python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_edge.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-finalize"
run() { # name, env...
  name=$1; shift
  env "$@" $B > gpurun_out/r8_$name.json 2> gpurun_out/r8_$name.err; echo rc=$?
  python -c "
import json;d=json.load(open('gpurun_out/r8_$name.json'));print('$name', '%.3e'%d['value'], round(d['ms_per_step'],2), d['roofline']['kernels_ms'])"
}
run tpb128 CITAM_PAIR_TPB=128
run tpb256 CITAM_PAIR_TPB=256
run l2persist CITAM_L2_PERSIST=1
run ilp1 CITAM_ACC_ILP=1
python scripts/r2_finalize_probe.py 30
