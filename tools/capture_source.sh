#!/bin/bash
# SASS-level stall attribution of one case (see capture_profiles.sh): the source page of an `ncu --set full` capture as CSV
set -u
out=gpurun_out; mkdir -p $out
case=$1; regex=$2; skip=$3
python tools/profile_cases.py $case > $out/plain_$case.log 2>&1 || { echo plain failed; exit 1; }
ncu --set full --import-source on --clock-control none -k regex:$regex -s $skip -c 1 -f -o $out/src_$case python tools/profile_cases.py $case > $out/ncu_$case.log 2>&1
ncu -i $out/src_$case.ncu-rep --page source --csv --print-source sass > $out/src_${case}_sass.csv 2>/dev/null
rm -f $out/src_$case.ncu-rep $out/ncu_$case.log
ls -la $out/src_${case}_sass.csv
