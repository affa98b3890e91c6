for c in 32 128 240; do
echo "== chunk $c"
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --chunk $c > gpurun_out/b3.json 2> gpurun_out/b3.err; echo rc=$?; python -c "
import json;d=json.load(open('gpurun_out/b3.json'));print(d['value'],d['roofline']['kernels_ms'],d['table_updates_per_contact_step'])"
done
