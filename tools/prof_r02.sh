set -x
NCU="ncu --set full --clock-control none --import-source on --launch-skip 3 -c 1"
$NCU -f -o gpurun_out/r02_gauss_r3_c3 python -m tools.prof_one gaussian 4320 7680 3 3 4 > gpurun_out/ncu_g3.log 2>&1
$NCU -f -o gpurun_out/r02_box_r3_c2 python -m tools.prof_one box 4096 4096 4 3 4 > gpurun_out/ncu_b3.log 2>&1
$NCU -f -o gpurun_out/r02_box_r16_c2 python -m tools.prof_one box 4096 4096 4 16 4 > gpurun_out/ncu_b16.log 2>&1
$NCU -f -o gpurun_out/r02_sobel_c3 python -m tools.prof_one sobel 4320 7680 3 1 4 > gpurun_out/ncu_s.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 6 -c 2 -f -o gpurun_out/r02_gauss_r15_c3 python -m tools.prof_one gaussian 4320 7680 3 15 4 > gpurun_out/ncu_g15.log 2>&1
ls -la gpurun_out/*.ncu-rep
