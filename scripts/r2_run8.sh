set -x
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-finalize"
for tpb in 128 256; do
CITAM_PAIR_TPB=$tpb $B > gpurun_out/r8_tpb$tpb.json 2> gpurun_out/r8_tpb$tpb.err; echo rc=$?
python -c "
import json;d=json.load(open('gpurun_out/r8_tpb$tpb.json'));print($tpb, '%.3e'%d['value'], d['ms_per_step'], d['roofline']['kernels_ms'])"
done
python scripts/r2_finalize_probe.py 30
