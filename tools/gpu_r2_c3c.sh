# round-2: one extra on a cubic track as an inline fast extra (XS = 3): tests, then C2 / C3 linear / C3 cubic timings
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -W always -x -k "chen25 or cubic or lmc or mw_lmc or stream or driver" ) > gpurun_out/c3c_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/c3c_pytest.log | tail -3
grep -n "^E  " gpurun_out/c3c_pytest.log | cut -c1-300 | head
( timeout 100 python tools/bench_k1.py 1000000; timeout 100 python tools/bench_k1.py 1000000 8 c3; timeout 200 python tools/bench_k1.py 1000000 8 c3cubic ) > gpurun_out/c3c.log 2>&1
grep -v "^+" gpurun_out/c3c.log | cut -c1-150
