# full GPU suite, default bench (both arms), launch list of one step (tools/prof_step.sh)
set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.log 2>&1; grep -E "passed|failed" gpurun_out/final_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/final_bench.log 2>&1; grep '^{' gpurun_out/final_bench.log | cut -c1-200
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/final_bench_ref.log 2>&1; grep '^{' gpurun_out/final_bench_ref.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash tools/prof_step.sh r02d
