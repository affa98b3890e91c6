CITAM_MERGE=1 timeout 900 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_edge.py tests/test_gpu_parity.py -x -q 2>&1 | tail -12
for w in 1 0; do
echo "== CITAM_MERGE=$w"
CITAM_MERGE=$w timeout 300 python bench.py --workload config2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/b2.json 2> gpurun_out/b2.err; echo rc=$?; python -c "
import json;d=json.load(open('gpurun_out/b2.json'));print(d['value'],d['e2e']['value'],d['roofline']['kernels_ms'],d['table_updates_per_contact_step'])"
CITAM_MERGE=$w timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/b3.json 2> gpurun_out/b3.err; echo rc=$?; tail -2 gpurun_out/b3.err | cut -c1-300; python -c "
import json;d=json.load(open('gpurun_out/b3.json'));print(d['value'],d['e2e']['value'],d['roofline']['kernels_ms'],d['table_updates_per_contact_step'])"
done
