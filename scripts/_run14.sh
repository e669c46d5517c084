mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_r2.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2.log
timeout 300 python scripts/run_configs.py > gpurun_out/run_configs_r2.txt 2>&1
awk '{print $1,$2,$3,$4,$5,$6,$7,$8,$9,$10,$11,$12,$13,$14}' gpurun_out/run_configs_r2.txt | sed -n '1,3p;11,13p'
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 100 --csv --log-file gpurun_out/launches_s0_r2.csv python scripts/s0_case.py > gpurun_out/s0_ncu.log 2>&1
tail -2 gpurun_out/s0_ncu.log
timeout 300 python bench.py --steps 2 --warmup 1 --no-s0 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('S1', d['value']/1e9, d['roofline']['frac'])"
