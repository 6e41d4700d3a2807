#!/bin/bash
# Round-2 validation: full GPU test suite, smoke, the default bench (Accel-101 headline + extras + cpu_baseline + parity),
# the reference arm, the per-frame interval parity of Accel-101 / Accel-18 / DFF.
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_final.log
tail -5 gpurun_out/r02_pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; tail -2 gpurun_out/r02_smoke_final.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final_101.json 2> gpurun_out/r02_bench_final_101.err; echo "bench rc $?"; tail -3 gpurun_out/r02_bench_final_101.err
for v in 101 18 dff; do timeout 600 python tools/interval_parity.py $v 2>&1 | tail -1 >> gpurun_out/r02_interval_parity.txt; done; cat gpurun_out/r02_interval_parity.txt
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc $?"; cat gpurun_out/r02_bench_reference.json | cut -c1-900
