# round-2 call X: shared-step kernel with the inline moving-Plummer term: tests + bench_c5
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -W always -x -k "restricted or nbody or perturber_set or c5" ) > gpurun_out/x_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/x_pytest.log | tail -3
grep -n "^E  " gpurun_out/x_pytest.log | cut -c1-300 | head
( timeout 400 python tools/bench_c5.py 10000000 100000 ) > gpurun_out/x_c5.log 2>&1
grep -v "^+" gpurun_out/x_c5.log | cut -c1-260
