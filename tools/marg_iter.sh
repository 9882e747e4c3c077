#!/bin/bash
# One gpurun call of the marg_schur_kernel tuning loop: parity tests, kernel-only durations (ncu, one wave of 148 windows per size), full capture of one size.
mkdir -p gpurun_out
python -m pytest tests/test_marg_schur.py -m gpu -q 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:marg_schur --csv --log-file gpurun_out/marg_times.csv python tools/marg_bench.py --windows ${WINDOWS:-296} --reps 1 > gpurun_out/marg_times.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/marg_times.csv')) if len(r) > 5 and r[0].isdigit()]
print('marg_schur_kernel durations (ms):', [round(float(r[-1].replace(',', '')) / (1e6 if 'ns' in r[-2] else 1e3 if 'us' in r[-2] else 1), 3) for r in rows])
PY
SZ=${1:-0}
if [ "$SZ" = "none" ]; then exit 0; fi
ncu --set full --clock-control none --import-source on -k regex:marg_schur -s 1 -c 1 -f -o gpurun_out/marg_$SZ python tools/marg_bench.py --windows 148 --reps 1 --only $SZ > gpurun_out/marg_ncu_$SZ.log 2>&1
ncu -i gpurun_out/marg_$SZ.ncu-rep --page source --csv > gpurun_out/marg_src_$SZ.csv
ncu -i gpurun_out/marg_$SZ.ncu-rep --page raw --csv > gpurun_out/marg_raw_$SZ.csv
