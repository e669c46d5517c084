mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 2 --warmup 1 --no-s0 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', '%.2f G/s frac %.4f rt_ms %.1f' % (d['value']/1e9, d['roofline']['frac'], d['phase_ms_per_step']['raytrace']), d['clocks']['sm_mhz'])"
}
run default X=1
for v in t512c1 t512c1f3 t256c2f3 t384c1 t384c1f2 t256c3i2 t320c2 t192c3 t128c4; do
  run $v C2B_LIB=gpurun_variants/libc2ray_b200_$v.so
done
run default_again X=1
run smem40 C2B_RT_SMEM_KB=40
run smem80 C2B_RT_SMEM_KB=80
run smem100 C2B_RT_SMEM_KB=100
run seg8 C2B_RT_SEGLEN=8
run seg16 C2B_RT_SEGLEN=16
run t512c1_smem200 C2B_LIB=gpurun_variants/libc2ray_b200_t512c1.so C2B_RT_SMEM_KB=200
run t512c1f3_smem200 C2B_LIB=gpurun_variants/libc2ray_b200_t512c1f3.so C2B_RT_SMEM_KB=200
run t384c1_smem200 C2B_LIB=gpurun_variants/libc2ray_b200_t384c1.so C2B_RT_SMEM_KB=200
