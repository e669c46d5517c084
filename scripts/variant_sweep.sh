#!/bin/bash
# correctness (cluster-only routing) and throughput of every work-group shape of the cluster kernel
for v in ${VARIANTS:-0 1 2 3}; do
  echo "== variant $v"
  C2B_CLUSTER_VARIANT=$v C2B_CLUSTER_MIN_NBOX=0 C2B_CLUSTER_MAX_SOURCES=1000000 C2B_DEBUG_CLUSTER=1 timeout 300 python scripts/gpu_check.py 2>&1 | grep -E "phih max|src1" | awk '{print $0}' | sort | uniq -c | sort -rn | head -3
  C2B_CLUSTER_VARIANT=$v C2B_CLUSTER_MAX_SOURCES=1000000 timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('variant $v', '%.2f G/s' % (d['value']/1e9), d['phase_ms_per_step'], d['updates_per_step'])"
done
