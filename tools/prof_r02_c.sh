set -x
# final round-2 pass: tests, smoke, both bench arms, launch list, captures of the kernels that changed last
python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; tail -n 2 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
python bench.py > gpurun_out/bench_final4.json 2> gpurun_out/bench_final4.err
python bench.py --impl reference > gpurun_out/bench_final4_ref.json 2> gpurun_out/bench_final4_ref.err
python -m tools.odd_pitch_bench > gpurun_out/r02_odd_pitch_final.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r02_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 6 -c 3 -f -o gpurun_out/r02_c4_kernels python -m tools.prof_c4 64 > gpurun_out/ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 3 -c 1 -f -o gpurun_out/r02_sobel_c1_odd python -m tools.prof_one sobel 2146 3239 3 1 4 > gpurun_out/ncu_s1.log 2>&1
tail -c 300 gpurun_out/bench_final4.json; tail -c 300 gpurun_out/bench_final4_ref.json; wc -l gpurun_out/r02_bench_launches_ncu.csv; ls -la gpurun_out/*.ncu-rep | tail -n 3
