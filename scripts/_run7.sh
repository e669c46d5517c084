mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 2 --warmup 1 --no-s0 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', '%.2f G/s frac %.4f rt_ms %.1f' % (d['value']/1e9, d['roofline']['frac'], d['phase_ms_per_step']['raytrace']), d['clocks']['sm_mhz'])"
}
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py tests/test_gpu_thermal.py -m gpu -x -q ) > gpurun_out/pytest_queue.log 2>&1
tail -4 gpurun_out/pytest_queue.log
run queue X=1
run noqueue C2B_LIB=gpurun_variants/libc2ray_b200_noq.so
run queue X=1
run noqueue C2B_LIB=gpurun_variants/libc2ray_b200_noq.so
