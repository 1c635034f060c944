timeout 120 python scripts/gpu_probe.py attn_f16 gemm_f16 2>&1 | cut -c1-420 | tail -20
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_r1_e.json 2> gpurun_out/bench_r1_e.err; tail -c 1900 gpurun_out/bench_r1_e.json; tail -3 gpurun_out/bench_r1_e.err
