set -x
python -m pytest tests/test_gpu_synthetic.py -m gpu -x -q 2>&1 | tail -5
