# round-2: production driver (batches in flight on their own streams, pinned staging of the results): tests + wall-clock per batch
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -W always -x -k "driver or impact" ) > gpurun_out/drv_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/drv_pytest.log | tail -3
grep -n "^E  " gpurun_out/drv_pytest.log | cut -c1-300 | head
( timeout 900 python tools/bench_driver.py 1000 500 8 0.004; timeout 900 python tools/bench_driver.py 5000 500 6 0.004 ) > gpurun_out/drv_bench3.log 2>&1
grep -v "^+" gpurun_out/drv_bench3.log | cut -c95-200
