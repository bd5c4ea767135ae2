# round-2: response kernel with the progenitor's moving sphere inline in the base force: tests (bit-identity incl. moving components), timings
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -W always -x -k "response or chen25 or generator or driver or c4" ) > gpurun_out/rp_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/rp_pytest.log | tail -3
grep -n "^E  " gpurun_out/rp_pytest.log | cut -c1-300 | head
( timeout 100 python tools/bench_response.py 10000 1000 1e-6; timeout 100 python tools/bench_response.py 10000 1000 1e-6 prog 0.1
  timeout 100 python tools/bench_response.py 2000 1000 1e-11; timeout 100 python tools/bench_response.py 2000 1000 1e-11 prog 0.1
  timeout 100 python tools/bench_response.py 100000 1000 1e-6; timeout 100 python tools/bench_response.py 100000 1000 1e-6 prog 0.1 ) > gpurun_out/rp.log 2>&1
grep -v "^+" gpurun_out/rp.log | grep "^C4\|^base" | cut -c1-130
