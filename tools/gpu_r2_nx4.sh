# round-2: fast extras with their rare fix-up inlined (no call site in the step loop): C3 linear / cubic timings, tests
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 100 python tools/bench_k1.py 1000000; timeout 100 python tools/bench_k1.py 1000000 8 c3; timeout 200 python tools/bench_k1.py 1000000 8 c3cubic ) > gpurun_out/nx4.log 2>&1
grep -v "^+" gpurun_out/nx4.log | cut -c1-150
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -W always -x -k "chen25 or cubic or lmc or mw_lmc or stream or c3 or host_entry" ) > gpurun_out/nx4_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/nx4_pytest.log | tail -3
grep -n "^E  " gpurun_out/nx4_pytest.log | cut -c1-300 | head
