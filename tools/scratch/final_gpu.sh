set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.log 2>&1; tail -5 gpurun_out/final_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/final_bench.log 2>&1; tail -4 gpurun_out/final_bench.log | cut -c1-300
