set -x
nvidia-smi -L
python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_edge.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15
for ch in 64 128; do
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-finalize --chunk $ch > gpurun_out/r2_split_c$ch.json 2> gpurun_out/r2_split_c$ch.err; echo rc=$?; tail -4 gpurun_out/r2_split_c$ch.err
done
CITAM_NO_SPLIT=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-finalize > gpurun_out/r2_nosplit.json 2> gpurun_out/r2_nosplit.err; echo rc=$?; tail -4 gpurun_out/r2_nosplit.err
