mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "warp" ) > gpurun_out/pytest_warp.log 2>&1
tail -5 gpurun_out/pytest_warp.log
timeout 300 python scripts/s0_case.py
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file gpurun_out/launches_s0.csv python scripts/s0_case.py > gpurun_out/s0_ncu.log 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('BENCH', d['value']/1e9, d['roofline']['frac'], d['S0'])"
