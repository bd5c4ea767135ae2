# round-2 call N: response kernel with round flags instead of phase barriers: timings first (tight timeouts), then the bit-identity test
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( for cfg in "10000 1000 1e-6" "2000 1000 1e-11" "100000 1000 1e-6"; do echo "$cfg"; timeout 60 python tools/bench_response.py $cfg; done ) > gpurun_out/n_response.log 2>&1
grep -v "^+" gpurun_out/n_response.log | cut -c1-150
( time timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -W always -x -k "response" ) > gpurun_out/n_pytest_resp.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/n_pytest_resp.log | tail
grep -n "^E  " gpurun_out/n_pytest_resp.log | cut -c1-300 | head -20
