# full GPU suite, default bench (both arms), launch list of one step and ncu --set full of the merged sub-pixel launch
set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.log 2>&1; tail -3 gpurun_out/final_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/final_bench.log 2>&1; tail -1 gpurun_out/final_bench.log | cut -c1-200
( time timeout 900 python bench.py --impl reference ) > gpurun_out/final_bench_ref.log 2>&1; tail -1 gpurun_out/final_bench_ref.log | cut -c1-300
bash tools/prof_step.sh r02c
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_halo_kernel -s 25 -c 2 -f -o gpurun_out/r02c_halo2 python tools/bench_layers.py imager > gpurun_out/ncu_r02c_halo2.log 2>&1; tail -2 gpurun_out/ncu_r02c_halo2.log
