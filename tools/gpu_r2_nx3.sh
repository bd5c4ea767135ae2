# round-2: shared-step kernel variants without an interpreter call site; saving kernel with the cooperative dense output inlined (variant build)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -W always -x -k "restricted or nbody or perturber_set or c5" ) > gpurun_out/nx3_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/nx3_pytest.log | tail -3
grep -n "^E  " gpurun_out/nx3_pytest.log | cut -c1-300 | head
( timeout 400 python tools/bench_c5.py 10000000 100000 ) > gpurun_out/nx3_c5.log 2>&1
grep -v "^+" gpurun_out/nx3_c5.log | grep "K5" | cut -c1-260
V=$GRAFT_REPO_ROOT/build/variants/coopinl.so
( for m in 64 16; do echo "call M=$m"; timeout 120 python tools/bench_snapshots.py 1000000 $m; echo "inline M=$m"; SSB_LIB_PATH=$V timeout 120 python tools/bench_snapshots.py 1000000 $m; done ) > gpurun_out/nx3_snap.log 2>&1
grep -v "^+" gpurun_out/nx3_snap.log | cut -c1-120
