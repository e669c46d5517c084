#!/bin/bash
# BASELINE configs[4] on one GPU: mesh x sources sweep of the S1 workload (one line per point)
mkdir -p gpurun_out
out=gpurun_out/sweep_r2_1gpu.txt
: > $out
for pt in "256 1000" "256 100000" "384 10000" "512 10000"; do
  set -- $pt
  timeout 400 python bench.py --mesh $1 --nsrc $2 --steps 1 --warmup 1 --no-s0 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('mesh %s^3 sources %s: %.1f G updates/s, %.3f s per evolve3D step (%d outer iterations, %.3g updates), ray tracer %.3f of the HBM roofline, chemistry %.1f ms per step' % ('$1', '$2', d['value']/1e9, d['ms_per_step']/1e3, d['outer_iterations_per_step'], d['updates_per_step'], d['roofline']['frac'], d['phase_ms_per_step']['chemistry']))" >> $out
  tail -1 $out
done
