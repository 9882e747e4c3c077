#!/bin/bash
# Timing + DRAM traffic of the kernels outside the solve step (one gpurun call; outputs under gpurun_out/).
set -x
mkdir -p gpurun_out
TAG=${1:-r2a}
python tools/aux_kernels.py 1024 150 > gpurun_out/aux_wall_$TAG.txt 2>&1
python tools/marg_bench.py --windows 296 --reps 2 > gpurun_out/marg_bench_$TAG.txt 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'preintegrate|marg|outlier|triangulate|shift_depth|projection_eval|imu_leg_eval|prior_eval|pack|unpermute' -c 80 --csv \
    --log-file gpurun_out/aux_launches_$TAG.csv python tools/aux_kernels.py 1024 150 > gpurun_out/aux_under_ncu_$TAG.log 2>&1
cat gpurun_out/aux_wall_$TAG.txt gpurun_out/marg_bench_$TAG.txt
