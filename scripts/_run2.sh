set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "warp" ) > gpurun_out/pytest_warp.log 2>&1
tail -15 gpurun_out/pytest_warp.log
timeout 600 python bench.py > gpurun_out/bench_warp.json 2> gpurun_out/bench_warp.err
tail -c 300 gpurun_out/bench_warp.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_warp.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['S0'], d.get('parity'))
PY
