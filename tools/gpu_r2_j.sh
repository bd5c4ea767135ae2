# round-2 call J: full GPU tests with the per-TU tableau operands and the production-driver tests, smoke, bench N=1
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1100 python -m pytest tests -m gpu -q -W always ) > gpurun_out/j_pytest_gpu.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/j_pytest_gpu.log | tail
grep -n "^E  " gpurun_out/j_pytest_gpu.log | cut -c1-300 | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/j_smoke.log 2>&1; tail -2 gpurun_out/j_smoke.log
timeout 400 python bench.py > gpurun_out/j_bench_n1.json 2> gpurun_out/j_bench_n1.err
cut -c1-900 gpurun_out/j_bench_n1.json
