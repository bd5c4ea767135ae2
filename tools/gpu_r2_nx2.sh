# round-2: saving orbit kernel without an extras path (XS = 4) vs the general one, same library; snapshot / dense tests
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( for m in 64 16 2; do echo "general M=$m"; SSB_ORBIT_NOEXTRAS=0 timeout 120 python tools/bench_snapshots.py 1000000 $m; echo "no-extras M=$m"; timeout 120 python tools/bench_snapshots.py 1000000 $m; done ) > gpurun_out/nx2.log 2>&1
grep -v "^+" gpurun_out/nx2.log | cut -c1-140
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -W always -x -k "snapshots or dense or ragged or failure or golden or printed" ) > gpurun_out/nx2_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/nx2_pytest.log | tail -3
grep -n "^E  " gpurun_out/nx2_pytest.log | cut -c1-300 | head
