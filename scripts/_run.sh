timeout 200 python scripts/gpu_probe.py attn_f16 2>&1 | cut -c1-300 | tail -8
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_r1_g.json 2> gpurun_out/bench_r1_g.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_g.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','clocks')})
print({k:d['roofline'][k] for k in ('achieved','frac','share_of_step')})
for e in d['roofline_other']: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in e.items() if k!='peak'})
PY
tail -3 gpurun_out/bench_r1_g.err
