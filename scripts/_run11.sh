mkdir -p gpurun_out
C2B_DEBUG_BALANCE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 300 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
txt=open('gpurun_out/bench_2gpu.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1])
print('2GPU value %.1f G/s ms/step %.1f frac %.3f'%(d['value']/1e9,d['ms_per_step'],d['roofline']['frac'])); print(d['phase_ms_per_step']); print(d['S0']); print(d['e2e']); print(d['config']['workload'][:200], d['updates_per_step'])
PY
