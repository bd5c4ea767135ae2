# round-2: progenitor step kernel without an extras path (XS = 4): stream timings at 1e6 and at one rank's share of an 8-way split, tests
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( for n in 1000000 125000; do echo "general n=$n"; SSB_ORBIT_NOEXTRAS=0 timeout 100 python tools/bench_k1.py $n; echo "no-extras n=$n"; timeout 100 python tools/bench_k1.py $n; done ) > gpurun_out/nx5.log 2>&1
grep -v "^+" gpurun_out/nx5.log | cut -c1-140
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -W always -x -k "stream or dense or golden or printed or c1 or c2 or pipeline or host_entry or chen25" ) > gpurun_out/nx5_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/nx5_pytest.log | tail -3
grep -n "^E  " gpurun_out/nx5_pytest.log | cut -c1-300 | head
