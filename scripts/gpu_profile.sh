#!/bin/bash
# ncu --set full captures of the hot kernels (one launch each), reports under gpurun_out/
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:bconv_tc -s 2 -c 1 -f -o gpurun_out/r1_bconv_tc python profiles/profile_bconv.py > gpurun_out/ncu1.log 2>&1
LSQ_C=64 LSQ_HW=56 timeout 300 $NCU -k regex:bconv_tc -s 2 -c 1 -f -o gpurun_out/r1_bconv_tc_c64 python profiles/profile_bconv.py > gpurun_out/ncu1b.log 2>&1
timeout 300 $NCU -k regex:solve_v1 -s 2 -c 1 -f -o gpurun_out/r1_solve_v1 python profiles/profile_solver.py > gpurun_out/ncu2.log 2>&1
timeout 300 $NCU -k regex:encode_act -s 2 -c 1 -f -o gpurun_out/r1_encode_act python profiles/profile_bconv.py > gpurun_out/ncu3.log 2>&1
LSQ_N=128 timeout 300 $NCU -k regex:stem_conv -s 2 -c 1 -f -o gpurun_out/r1_stem_conv python profiles/profile_stem.py > gpurun_out/ncu4.log 2>&1
for f in gpurun_out/ncu*.log; do tail -n 2 $f; done
