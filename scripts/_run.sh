timeout 300 python scripts/gpu_probe.py attn 2>&1 | cut -c1-300 | tail -8
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_r1_d.json 2> gpurun_out/bench_r1_d.err; tail -c 1700 gpurun_out/bench_r1_d.json; tail -3 gpurun_out/bench_r1_d.err
