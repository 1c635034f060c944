timeout 600 python -m pytest tests/test_fusion_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -15
