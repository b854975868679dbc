#!/bin/bash
# Final measurement job of round 2 (one gpurun call, 1 GPU): full GPU suite, bench lines of every configuration,
# the reference arm, launch lists and the ncu --set full capture of the dominant kernel.  Outputs: gpurun_out/r02e_*.
set -u
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r02e_gputests.log
timeout 400 python bench.py > $O/r02e_bench_cfg3.json 2> $O/r02e_bench_cfg3.err
for c in 2 4 5; do timeout 500 python bench.py --config $c --steps 50 > $O/r02e_bench_cfg$c.json 2> $O/r02e_bench_cfg$c.err; done
timeout 300 python bench.py --config 4 --reg tv3d --steps 30 --no-cpu-baseline > $O/r02e_bench_cfg4_tv3d.json 2>/dev/null
timeout 300 python bench.py --impl reference > $O/r02e_bench_reference_arm.json 2>/dev/null
K='k_tile|k_band|k_btv|k_finish|k_tv3d|k_stage|k_reg|k_cg|k_lbfgs'
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 40 --csv --log-file $O/r02e_launches_cfg3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --solve-iters 0 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 60 --csv --log-file $O/r02e_launches_cfg5.csv python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline --solve-iters 0 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 60 --csv --log-file $O/r02e_launches_cfg4.csv python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline --solve-iters 0 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile_zt -s 3 -c 1 -o $O/r02e_k_tile_zt python bench.py --steps 2 --warmup 3 --no-cpu-baseline --solve-iters 0 > $O/r02e_ncu.log 2>&1
cat $O/r02e_gputests.log
python tools/launch_table.py $O/r02e_launches_cfg3.csv $O/r02e_launches_cfg5.csv $O/r02e_launches_cfg4.csv
for f in cfg3 cfg2 cfg4 cfg5 cfg4_tv3d; do python - <<PY
import json
try:
    d=json.loads(open("$O/r02e_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "step %.4f ms" % d["ms_per_step"], "kernel %.4f" % d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"], "e2e %.3f" % d["e2e"]["ms_per_step"], "solve/it %.3f" % d["solve"]["ms_per_iteration"], "clocks", d["clocks"])
except Exception as e:
    print("$f", "FAILED", e)
PY
done
