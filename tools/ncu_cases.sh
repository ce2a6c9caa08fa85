#!/bin/bash
# ncu --set full captures of single conv layers (second repetition), small reports + raw csv pages
set -u
out=gpurun_out/ncu_cases; mkdir -p $out
K='regex:edge_|thin|conv_tma_tc|conv_wgrad_tma|pixel_reduce|gather_gemm'
run() { name=$1; shift
  timeout 120 python tools/tc_case.py "$@" 2 > $out/$name.plain.log 2>&1 || { echo "$name plain run failed"; return; }
  timeout 400 ncu --set full --import-source on --clock-control none -k "$K" --launch-skip 3 -c 3 -o $out/$name -f python tools/tc_case.py "$@" 2 > $out/$name.ncu.log 2>&1
  ncu -i $out/$name.ncu-rep --page raw --csv > $out/$name.raw.csv 2>/dev/null
  ls -la $out/$name.ncu-rep
}
run dfirst 0 16 2 32 512 512 4 2 2
run head   0 16 256 1 66 66 4 1 2
run l4     0 16 128 256 65 65 4 1 2
du -sh gpurun_out
