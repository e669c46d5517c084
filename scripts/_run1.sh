set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_r2.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2.log
timeout 600 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
tail -c 600 gpurun_out/bench_r2.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-s0 > gpurun_out/b_ncu1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:raytrace_kernel -s 4 -c 1 -f -o gpurun_out/rt_r2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-s0 > gpurun_out/b_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chemistry_kernel -s 3 -c 1 -f -o gpurun_out/chem_r2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-s0 > gpurun_out/b_ncu3.log 2>&1
timeout 300 python scripts/run_configs.py > gpurun_out/run_configs_r2.txt 2>&1
ls -la gpurun_out
