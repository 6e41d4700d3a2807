#!/bin/bash
# ncu --set full of the deformable im2col kernel, per-thread (ACCEL_DCN_TILED=0) and tiled variants, inside the Accel-101 interval plan
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
for v in 0 1; do
ACCEL_DCN_TILED=$v timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:dcn_col -s 15 -c 2 -o gpurun_out/r02_dcn_col_v$v -f \
  python tools/interval_times.py --version 101 > gpurun_out/r02_ncu_dcn_$v.log 2>&1; echo "rc $?"
done
ls -la gpurun_out/r02_dcn_col_v*.ncu-rep
