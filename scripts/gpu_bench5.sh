for c in 29 30 31; do
echo "== pair-cap-log2 $c"
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --pair-cap-log2 $c > gpurun_out/b3.json 2> gpurun_out/b3.err; echo rc=$?; python -c "
import json;d=json.load(open('gpurun_out/b3.json'));print(d['value'],d['roofline']['kernels_ms'])"
done
