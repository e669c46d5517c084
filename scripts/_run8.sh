mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rA ) > gpurun_out/pytest_gpu_multi_2gpu_r2.log 2>&1
tail -8 gpurun_out/pytest_gpu_multi_2gpu_r2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 400 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
txt=open('gpurun_out/bench_2gpu.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1])
print('2GPU value %.1f G/s ms/step %.1f frac %.3f'%(d['value']/1e9,d['ms_per_step'],d['roofline']['frac'])); print(d['phase_ms_per_step']); print(d['S0']); print(d['e2e'])
PY
C2B_NO_BALANCE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-s0 --no-e2e 2>/dev/null | python -c "
import json,sys
txt=sys.stdin.read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1])
print('NO_BALANCE value %.1f G/s ms/step %.1f'%(d['value']/1e9,d['ms_per_step'])); print(d['phase_ms_per_step'])"
