# round-2 call S: nvcc -split-compile 0 (parallel split of each translation unit) vs 1 (whole-unit optimisation): same source, same box
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$GRAFT_REPO_ROOT/build/variants/nosplit.so
( for rep in 1 2; do echo "split-compile 0"; timeout 100 python tools/bench_k1.py 1000000; echo "split-compile 1"; SSB_LIB_PATH=$V timeout 100 python tools/bench_k1.py 1000000; done
  for cfg in "10000 1000 1e-6" "2000 1000 1e-11" "100000 1000 1e-6"; do echo "split-compile 0: $cfg"; timeout 100 python tools/bench_response.py $cfg; echo "split-compile 1: $cfg"; SSB_LIB_PATH=$V timeout 100 python tools/bench_response.py $cfg; done
  echo "split-compile 0 c3"; timeout 100 python tools/bench_k1.py 1000000 8 c3; echo "split-compile 1 c3"; SSB_LIB_PATH=$V timeout 100 python tools/bench_k1.py 1000000 8 c3
  echo "split-compile 0 dopri5"; timeout 100 python tools/bench_k1.py 1000000 5; echo "split-compile 1 dopri5"; SSB_LIB_PATH=$V timeout 100 python tools/bench_k1.py 1000000 5
  for m in 64 16; do echo "split-compile 0 snapshots"; timeout 120 python tools/bench_snapshots.py 1000000 $m; echo "split-compile 1 snapshots"; SSB_LIB_PATH=$V timeout 120 python tools/bench_snapshots.py 1000000 $m; done ) > gpurun_out/s_split.log 2>&1
grep -v "^+" gpurun_out/s_split.log | cut -c1-150
