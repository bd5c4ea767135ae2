# round-2 call W: shared-step kernel (K5) with persistent CTAs: tests, then bench_c5 with 3 and 2 CTAs per SM
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -W always -x -k "restricted or nbody or perturber_set or growing" ) > gpurun_out/w_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/w_pytest.log | tail -3
grep -n "^E  " gpurun_out/w_pytest.log | cut -c1-300 | head
( echo "3 CTAs/SM"; timeout 400 python tools/bench_c5.py 10000000 100000; echo "2 CTAs/SM"; SSB_LIB_PATH=$GRAFT_REPO_ROOT/build/variants/k5_2cta.so timeout 400 python tools/bench_c5.py 10000000 100000 ) > gpurun_out/w_c5.log 2>&1
grep -v "^+" gpurun_out/w_c5.log | cut -c1-260
