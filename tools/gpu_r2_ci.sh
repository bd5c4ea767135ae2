# round-2: saving kernel with the cooperative dense output inlined vs out of line
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$GRAFT_REPO_ROOT/build/variants/coopinl.so
( for m in 64 16 2; do echo "call M=$m"; timeout 120 python tools/bench_snapshots.py 1000000 $m; echo "inline M=$m"; SSB_LIB_PATH=$V timeout 120 python tools/bench_snapshots.py 1000000 $m; done ) > gpurun_out/ci_snap.log 2>&1
grep -v "^+" gpurun_out/ci_snap.log | cut -c1-150
