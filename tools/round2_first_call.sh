#!/bin/bash
# First GPU job of round 2 (one `gpurun --timeout 1500 -- bash tools/round2_first_call.sh`): everything
# that was written at the end of round 1 after the GPU budget was spent.
#   1. the gated tests of the device-resident solver (CUDA vector backend never ran on a device)
#   2. the whole GPU suite with k_tile_z as the default
#   3. A/B k_tile vs k_tile_z, then the bench line, launch list and ncu --set full of k_tile_z
# Everything lands in gpurun_out/; summarise into profiles/r02_* with tools/ncu_summary.py and
# tools/phase_breakdown.py, and put the k_tile_z DRAM bytes into profiles/traffic.json.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(SRB_RUN_PENDING=1 timeout 300 python -m pytest tests/test_gpu_cg.py -q -x -s 2>&1 | tail -40) > gpurun_out/r2_cg_tests.log
(SRB_RUN_PENDING=1 timeout 300 python -m pytest tests/test_gpu_zlayout.py -q -s 2>&1 | tail -20) > gpurun_out/r2_zholes_tests.log
(timeout 900 python -m pytest tests -q -x -m gpu --durations=15 2>&1 | tail -40) > gpurun_out/r2_gpu_tests.log
(timeout 120 python tools/ab_zlayout.py 60 2>&1) > gpurun_out/r2_ab_zlayout.log
(timeout 300 python bench.py 2>gpurun_out/r2_bench.err) > gpurun_out/r2_bench.json
(timeout 300 python bench.py --steps 20 --no-cpu-baseline --solve-iters 20 2>gpurun_out/r2_bench_solve.err) > gpurun_out/r2_bench_solve.json
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline) > gpurun_out/r2_ncu_launch.log 2>&1
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_z -s 3 -c 1 \
    -o gpurun_out/r2_k_tile_z python bench.py --steps 2 --warmup 3 --no-cpu-baseline) > gpurun_out/r2_ncu_full.log 2>&1
tail -5 gpurun_out/r2_cg_tests.log gpurun_out/r2_zholes_tests.log gpurun_out/r2_gpu_tests.log
cat gpurun_out/r2_ab_zlayout.log | head -30
cat gpurun_out/r2_bench.json gpurun_out/r2_bench_solve.json
