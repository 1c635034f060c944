timeout 600 python -m pytest tests/test_golden_gpu.py -m gpu -x -q -k logmel 2>&1 | tail -8
