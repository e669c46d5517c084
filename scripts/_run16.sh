mkdir -p gpurun_out
N=$1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --mesh 512 --nsrc 100000 --scaling strong --steps 2 --warmup 1 --no-s0 > gpurun_out/bench_512_${N}gpu.json 2> gpurun_out/bench_512_${N}gpu.err
python - $N <<'PY'
import json,sys
txt=open('gpurun_out/bench_512_%sgpu.json'%sys.argv[1]).read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1])
print(sys.argv[1],'GPU value %.1f G/s ms/step %.1f frac %.3f'%(d['value']/1e9,d['ms_per_step'],d['roofline']['frac'])); print(d['phase_ms_per_step']); print(d['e2e']); print(d['clocks'])
PY
