# final evidence of a round on one GPU: the whole -m gpu suite, smoke, the default bench line, the other
# workloads' lines, the reference arm
set -x
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/r2_bench_config3.json 2> gpurun_out/r2_bench_config3.err; echo rc=$?; tail -3 gpurun_out/r2_bench_config3.err | cut -c1-400
python bench.py --workload config2 --steps 6 --warmup 3 > gpurun_out/r2_bench_config2.json 2> gpurun_out/r2_bench_config2.err; echo rc=$?
python bench.py --workload config5 --steps 3 --warmup 3 > gpurun_out/r2_bench_config5.json 2> gpurun_out/r2_bench_config5.err; echo rc=$?; tail -3 gpurun_out/r2_bench_config5.err | cut -c1-400
python bench.py --mode blocks --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_config3_blocks.json 2> gpurun_out/r2_bench_config3_blocks.err; echo rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo rc=$?
python - <<'PY'
import json
for f in ["config3","config2","config5","config3_blocks","ref"]:
    try:
        d=json.load(open("gpurun_out/r2_bench_%s.json"%f))
        print(f, "value %.3e e2e %.3e ms/step %.2f" % (d["value"], (d.get("e2e") or {}).get("value",0), d["ms_per_step"]), "cpu", (d.get("cpu_baseline") or {}).get("value"), "clocks", d.get("clocks"))
        if d.get("roofline"): print("   roof", {k:d["roofline"][k] for k in ("kernel","achieved","frac","traffic","whole_step_frac","kernels_ms")}, d["roofline"].get("random_update_ceiling"))
        print("   ", {k:d.get(k) for k in ("contacts_per_agent_step","table_updates_per_contact_step","binned_fraction_of_agent_steps","active_fraction","finalize","e2e_with_results","day")})
    except Exception as e: print(f, "ERR", e)
PY
