mkdir -p gpurun_out
C2B_DEBUG_BALANCE=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 --no-s0 --no-e2e > gpurun_out/bench_8gpu_weak_balance.json 2> gpurun_out/bench_8gpu_weak_balance.err
grep -c "dealt again" gpurun_out/bench_8gpu_weak_balance.err; grep "dealt again" gpurun_out/bench_8gpu_weak_balance.err | tail -4
C2B_NO_BALANCE=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 3 --no-s0 --no-e2e > gpurun_out/bench_8gpu_weak_nobalance.json 2> gpurun_out/bench_8gpu_weak_nobalance.err
for f in balance nobalance; do python - $f <<'PY'
import json,sys
txt=open('gpurun_out/bench_8gpu_weak_%s.json'%sys.argv[1]).read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1])
print(sys.argv[1], 'value %.1f G/s ms/step %.1f frac %.3f'%(d['value']/1e9,d['ms_per_step'],d['roofline']['frac'])); print(d['phase_ms_per_step']); print(d['clocks'])
PY
done
