timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_r1_g.json 2> gpurun_out/bench_r1_g.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_g.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','clocks')})
print(json.dumps(d['roofline'])[:600])
for e in d['roofline_other']: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in e.items()})
PY
tail -3 gpurun_out/bench_r1_g.err
