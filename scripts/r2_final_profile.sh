set -x
python bench.py > gpurun_out/r2_bench_config3.json 2> gpurun_out/r2_bench_config3.err; echo rc=$?
python -c "
import json;d=json.load(open('gpurun_out/r2_bench_config3.json'));print('%.3e e2e %.3e'%(d['value'],d['e2e']['value']), round(d['ms_per_step'],2), d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['kernels_ms'], d['roofline'].get('random_update_ceiling',{}).get('frac_of_ceiling'), d['cpu_baseline']['value'], d['e2e_with_results'])"
bash scripts/r2_profile.sh
