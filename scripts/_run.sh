timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_r1_h.json 2> gpurun_out/bench_r1_h.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_h.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','clocks','gpu_launches')})
print({k:d['roofline'][k] for k in ('achieved','frac','share_of_step')})
for e in d['roofline_other']: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in e.items() if k in ('kernel','achieved','frac','share_of_step')})
print(d['cpu_baseline'])
PY
tail -3 gpurun_out/bench_r1_h.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
