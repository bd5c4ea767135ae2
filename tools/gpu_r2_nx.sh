# round-2: final-state orbit kernel without any extras path in its step loop (XS = 4) vs the general kernel, same library
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( for rep in 1 2 3; do echo "general"; SSB_ORBIT_NOEXTRAS=0 timeout 100 python tools/bench_k1.py 1000000; echo "no-extras"; timeout 100 python tools/bench_k1.py 1000000; done
  echo "general dopri5"; SSB_ORBIT_NOEXTRAS=0 timeout 100 python tools/bench_k1.py 1000000 5; echo "no-extras dopri5"; timeout 100 python tools/bench_k1.py 1000000 5 ) > gpurun_out/nx.log 2>&1
grep -v "^+" gpurun_out/nx.log | cut -c1-140
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_adaptive_parity.py -m gpu -q -W always -x -k "stream or orbit or golden or printed or c1 or c2 or pipeline or adaptive or lock or ensemble or host_entry" ) > gpurun_out/nx_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/nx_pytest.log | tail -3
grep -n "^E  " gpurun_out/nx_pytest.log | cut -c1-300 | head
