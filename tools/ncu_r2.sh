#!/bin/bash
# round-2 ncu captures (one GPU): --set full of the dominant kernels on the layers that dominate the fcgan step
set -x
mkdir -p gpurun_out/ncu_r2
N="ncu --set full --clock-control none --import-source on -f"
SGK_PATCH=0 $N -k regex:conv_tma_tc_kernel --launch-skip 2 -c 2 -o gpurun_out/ncu_r2/l4_tile python tools/layer_bench.py C128-256_65_N16 > gpurun_out/ncu_r2/l4_tile.log 2>&1
$N -k regex:conv_wgrad_tma_kernel --launch-skip 1 -c 1 -o gpurun_out/ncu_r2/l4_wgrad python tools/layer_bench.py C128-256_65_N16 > gpurun_out/ncu_r2/l4_wgrad.log 2>&1
SGK_PATCH=2 $N -k regex:conv_patch_tc_kernel --launch-skip 2 -c 2 -o gpurun_out/ncu_r2/l2_patch python tools/layer_bench.py C32-64_257_N16 > gpurun_out/ncu_r2/l2_patch.log 2>&1
$N -k regex:conv_wgrad_tma_kernel --launch-skip 1 -c 1 -o gpurun_out/ncu_r2/l2_wgrad python tools/layer_bench.py C32-64_257_N16 > gpurun_out/ncu_r2/l2_wgrad.log 2>&1
$N -k regex:norm_fused --launch-skip 0 -c 2 -o gpurun_out/ncu_r2/norm_fused python tools/norm_bench.py > gpurun_out/ncu_r2/norm.log 2>&1
ls -la gpurun_out/ncu_r2
