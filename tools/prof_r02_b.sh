set -x
# per-frame traffic of the three headline kernels (launches 6,7,8 = third round of gaussian, box, sobel)
ncu --set full --clock-control none --import-source on --launch-skip 6 -c 3 -f -o gpurun_out/r02_c4_kernels python -m tools.prof_c4 64 > gpurun_out/ncu_c4.log 2>&1
# the bench command's launch list (one-pass duration metric)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r02_bench_under_ncu.log 2>&1
tail -3 gpurun_out/ncu_c4.log; wc -l gpurun_out/r02_bench_launches_ncu.csv
