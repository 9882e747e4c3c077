#!/bin/bash
# Everything the round-end evidence needs, in one gpurun call (outputs under gpurun_out/; tools/make_profiles.py turns them into profiles/).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/gpu_tests.txt
CERB_TRACE=1 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
make -s prof && python tools/phase_profile.py 148 150 1 > gpurun_out/phase_final.txt 2>&1
python tools/marg_phase.py 148 > gpurun_out/marg_phase_final.txt 2>&1
WINDOWS=296 bash tools/marg_iter.sh none > gpurun_out/marg_kernel_times_final.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vilo_solve -c 1 -o gpurun_out/solve_final_1024 -f python tools/profile_solve.py 1024 150 1 > gpurun_out/ncu_final_1024.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vilo_solve -c 1 -o gpurun_out/solve_final_148 -f python tools/profile_solve.py 148 150 1 > gpurun_out/ncu_final_148.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:marg_schur -s 1 -c 1 -o gpurun_out/marg_schur_final -f python tools/aux_kernels.py 1024 150 > gpurun_out/ncu_marg.log 2>&1
bash tools/gpu_aux_profile.sh final > gpurun_out/aux_final.log 2>&1
cat gpurun_out/gpu_tests.txt
cut -c1-260 gpurun_out/bench_final.json
cut -c1-260 gpurun_out/bench_reference.json
tail -3 gpurun_out/aux_wall_final.txt
