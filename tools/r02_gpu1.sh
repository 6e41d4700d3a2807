#!/bin/bash
# round 2, first GPU pass: new tests (interval plan, config 1, full-size oracle parity), then the default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_interval.py tests/test_gpu_config1.py -x -q -s > gpurun_out/r02_t1.log 2>&1; echo "exit $?" >> gpurun_out/r02_t1.log
tail -25 gpurun_out/r02_t1.log
timeout 900 python -m pytest tests/test_gpu_fullsize_oracle.py -x -q -s > gpurun_out/r02_t2.log 2>&1; echo "exit $?" >> gpurun_out/r02_t2.log
tail -12 gpurun_out/r02_t2.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench exit $?"
tail -c 3000 gpurun_out/r02_bench_a.json; tail -5 gpurun_out/r02_bench_a.err
