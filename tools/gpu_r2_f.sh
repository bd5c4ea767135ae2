# round-2 call F: GPU tests (perturber set, growing potential, retired-items fixes), C5 end-to-end tool
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -W always ) > gpurun_out/f_pytest_gpu.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/f_pytest_gpu.log | tail
grep -n "^E  " gpurun_out/f_pytest_gpu.log | cut -c1-300 | head -20
timeout 600 python tools/bench_c5.py 1e7 1e6 > gpurun_out/f_c5.log 2>&1
cat gpurun_out/f_c5.log | tail -12
