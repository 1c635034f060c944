timeout 600 python -m pytest tests/test_vit_gpu.py -m gpu -x -q 2>&1 | tail -15
