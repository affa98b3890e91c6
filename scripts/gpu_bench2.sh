for g in 0 1 2 3; do
echo "== variant $g"
CITAM_VARIANT=$g timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b3.json 2> gpurun_out/b3.err; echo rc=$?; python -c "
import json;d=json.load(open('gpurun_out/b3.json'));print(d['value'],d['roofline']['kernels_ms'])"
done
CITAM_VARIANT=3 timeout 600 python -m pytest tests/test_gpu_synthetic.py -x -q 2>&1 | tail -2
