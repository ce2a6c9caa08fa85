#!/bin/bash
# usage: tools/ncu_one.sh <name> <kernel-regex> <skip> <count> -- tc_case args...   (second repetition of tools/tc_case.py)
set -u
name=$1; regex=$2; skip=$3; count=$4; shift 5
out=gpurun_out/ncu_cases; mkdir -p $out
timeout 120 python tools/tc_case.py "$@" 2 > $out/$name.plain.log 2>&1 || { echo "$name plain run failed"; tail -3 $out/$name.plain.log; exit 1; }
timeout 400 ncu --set full --import-source on --clock-control none -k "regex:$regex" --launch-skip $skip -c $count -o $out/$name -f python tools/tc_case.py "$@" 2 > $out/$name.ncu.log 2>&1
ncu -i $out/$name.ncu-rep --page raw --csv > $out/$name.raw.csv 2>/dev/null
ncu -i $out/$name.ncu-rep --page source --csv --print-source sass > $out/$name.sass.csv 2>/dev/null
ls -la $out/$name.ncu-rep
