#!/bin/bash
# Round-end validation: full GPU test suite, smoke, default bench (with cpu_baseline), other configs, launch list.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_final.log
tail -5 gpurun_out/pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -2 gpurun_out/smoke_final.log
timeout 900 python bench.py > gpurun_out/bench_final_dff.json 2> gpurun_out/bench_final_dff.err; cat gpurun_out/bench_final_dff.json
for v in 18 34 50 101; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --version $v > gpurun_out/bench_final_$v.json 2> gpurun_out/bench_final_$v.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_dff_final.csv \
    python tools/profile_step.py --version dff --intervals 2 --flags 2 > gpurun_out/ncu_launches_final.log 2>&1
