#!/bin/bash
# Regenerates the single-GPU evidence under profiles/ (run on a B200 box): smoke, pytest -m gpu, bench (both arms),
# launch lists, ncu --set full captures of the three hot kernels, configs 1/2, compute-sanitizer.
# Outputs go to gpurun_out/; summarise the .ncu-rep files with scripts/ncu_summary.py / scripts/ncu_lines.py.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_r2.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2.log
timeout 600 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
tail -c 300 gpurun_out/bench_r2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2>/dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-s0 > gpurun_out/b_ncu1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:raytrace_kernel -s 4 -c 1 -f -o gpurun_out/rt_r2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-s0 > gpurun_out/b_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chemistry_kernel -s 3 -c 1 -f -o gpurun_out/chem_r2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-s0 > gpurun_out/b_ncu3.log 2>&1
timeout 300 python scripts/s0_case.py > gpurun_out/s0_case.log 2>&1; tail -2 gpurun_out/s0_case.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 100 --csv --log-file gpurun_out/launches_s0_r2.csv python scripts/s0_case.py > gpurun_out/s0_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:raytrace_warp_kernel -s 45 -c 1 -f -o gpurun_out/warp_r2 python scripts/s0_case.py > gpurun_out/s0_ncu2.log 2>&1
timeout 300 python scripts/run_configs.py > gpurun_out/run_configs_r2.txt 2>&1
( timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_case.py 2>&1 | grep -E "^(cta|cluster|warp)|ERROR SUMMARY|Invalid|Error" | head -40 ) > gpurun_out/sanitizer_r2.txt
( timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_case.py 2>&1 | grep -E "^(cta|cluster|warp)|RACECHECK SUMMARY|hazard|Error" | head -40 ) >> gpurun_out/sanitizer_r2.txt
cat gpurun_out/sanitizer_r2.txt
ls -la gpurun_out
