mkdir -p gpurun_out
C2B_DEBUG_BALANCE=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_8gpu_weak.json 2> gpurun_out/bench_8gpu_weak.err
grep -c "dealt again" gpurun_out/bench_8gpu_weak.err; grep "dealt again" gpurun_out/bench_8gpu_weak.err | tail -3
python - <<'PY'
import json
txt=open('gpurun_out/bench_8gpu_weak.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1])
print('8GPU value %.1f G/s ms/step %.1f frac %.3f'%(d['value']/1e9,d['ms_per_step'],d['roofline']['frac'])); print(d['phase_ms_per_step']); print(d['S0']); print(d['e2e']); print(d['updates_per_step'], d['clocks'])
PY
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 3 --warmup 3 --no-s0 > gpurun_out/bench_4gpu_weak.json 2> gpurun_out/bench_4gpu_weak.err
python - <<'PY'
import json
txt=open('gpurun_out/bench_4gpu_weak.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1])
print('4GPU value %.1f G/s ms/step %.1f frac %.3f'%(d['value']/1e9,d['ms_per_step'],d['roofline']['frac'])); print(d['phase_ms_per_step']); print(d['e2e'])
PY
