# round-2 call Q: rotating bars (jets) - parity tests, full GPU suite, and a regression check of the hot kernels' timings
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -W always -x -k "rotating_bars or closed_forms or third" ) > gpurun_out/q_pytest_bars.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/q_pytest_bars.log | tail -3
grep -n "^E  " gpurun_out/q_pytest_bars.log | cut -c1-300 | head -20
( timeout 100 python tools/bench_k1.py 1000000; timeout 100 python tools/bench_k1.py 1000000 8 c3; timeout 100 python tools/bench_response.py 10000 1000 1e-6; timeout 100 python tools/bench_response.py 2000 1000 1e-11
  timeout 120 python tools/bench_snapshots.py 1000000 64 ) > gpurun_out/q_regress.log 2>&1
grep -v "^+" gpurun_out/q_regress.log | cut -c1-200
( time timeout 1100 python -m pytest tests -m gpu -q -W always ) > gpurun_out/q_pytest_gpu.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/q_pytest_gpu.log | tail
