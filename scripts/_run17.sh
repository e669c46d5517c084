mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_r2.log 2>&1
tail -4 gpurun_out/pytest_gpu_r2.log
timeout 600 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
tail -c 200 gpurun_out/bench_r2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2.json').read().strip().splitlines()[-1])
print(d['value']/1e9, d['roofline']['frac'], d['S0']['value']/1e9, d['e2e']['value']/1e9, d['parity'], d['cpu_baseline']['cores'])
PY
