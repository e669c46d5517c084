set -x
mkdir -p gpurun_out
for i in 1 2; do
( cd gpurun_variants/prev_tree && timeout 300 python bench.py --steps 3 --warmup 2 --no-s0 --no-e2e --no-cpu-baseline ) 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PREV', d['value']/1e9, d['roofline']['frac'], d['clocks'])"
timeout 300 python bench.py --steps 3 --warmup 2 --no-s0 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NEW ', d['value']/1e9, d['roofline']['frac'], d['clocks'])"
done
timeout 300 python scripts/s0_case.py
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 120 --csv --log-file gpurun_out/launches_s0.csv python scripts/s0_case.py > gpurun_out/s0_ncu.log 2>&1
tail -3 gpurun_out/s0_ncu.log
