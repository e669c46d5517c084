mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rA ) > gpurun_out/pytest_gpu_multi_2gpu_r2.log 2>&1
tail -4 gpurun_out/pytest_gpu_multi_2gpu_r2.log
C2B_DEBUG_BALANCE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-s0 --no-e2e > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
grep -c "dealt again" gpurun_out/bench_2gpu.err; grep "dealt again" gpurun_out/bench_2gpu.err | tail -3
python - <<'PY'
import json
txt=open('gpurun_out/bench_2gpu.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1])
print('2GPU value %.1f G/s ms/step %.1f frac %.3f'%(d['value']/1e9,d['ms_per_step'],d['roofline']['frac'])); print(d['phase_ms_per_step'])
PY
