# round-2: item force with one code path for the mass and radius blocks: response tests + timings
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -W always -x -k "response or chen25 or generator or driver or c4 or edge" ) > gpurun_out/uf_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/uf_pytest.log | tail -3
grep -n "^E  " gpurun_out/uf_pytest.log | cut -c1-300 | head
( for rep in 1 2; do timeout 100 python tools/bench_response.py 10000 1000 1e-6; done; timeout 100 python tools/bench_response.py 2000 1000 1e-11; timeout 100 python tools/bench_response.py 100000 1000 1e-6 ) > gpurun_out/uf.log 2>&1
grep -v "^+" gpurun_out/uf.log | grep "^C4" | cut -c1-130
