# round-2 call Z4: full GPU suite on the final build (moving-progenitor response test included)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1100 python -m pytest tests -m gpu -q -W always ) > gpurun_out/z4_pytest_gpu.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/z4_pytest_gpu.log | tail
grep -n "^E  " gpurun_out/z4_pytest_gpu.log | cut -c1-300 | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z4_smoke.log 2>&1; tail -1 gpurun_out/z4_smoke.log
timeout 500 python bench.py > gpurun_out/z4_bench_n1.json 2> gpurun_out/z4_bench_n1.err; cut -c1-400 gpurun_out/z4_bench_n1.json
