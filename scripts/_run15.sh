timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "other_subbox" 2>&1 | tail -3
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 2 --warmup 1 --no-s0 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', '%.2f G/s frac %.4f' % (d['value']/1e9, d['roofline']['frac']))"
}
run default X=1
run seg4 C2B_RT_SEGLEN=4
run seg6 C2B_RT_SEGLEN=6
run seg10 C2B_RT_SEGLEN=10
run default X=1
