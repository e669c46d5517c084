mkdir -p gpurun_out
echo "== pair (default)"; timeout 300 python scripts/s0_case.py | tail -1
echo "== sequential"; C2B_LIB=gpurun_variants/libc2ray_b200_chemseq.so timeout 300 python scripts/s0_case.py | tail -1
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_thermal.py -m gpu -x -q -k "auto or thermal" ) > gpurun_out/pytest_chem.log 2>&1
tail -5 gpurun_out/pytest_chem.log
C2B_LIB=gpurun_variants/libc2ray_b200_chemseq.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "auto" 2>&1 | tail -3
