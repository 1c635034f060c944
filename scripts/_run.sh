timeout 120 python scripts/gpu_probe.py gemm_f16 2>&1 | cut -c1-330 | tail -11
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_r1_f.json 2> gpurun_out/bench_r1_f.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_f.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','clocks','roofline','roofline_other')})
PY
