# round-2 call O: CTA-aligned step loops (one CTA-wide vote per iteration) A/B: final-state stream kernel and saving kernel
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$GRAFT_REPO_ROOT/build/variants/align3.so
( for rep in 1 2; do echo default; timeout 100 python tools/bench_k1.py 1000000; echo aligned; SSB_LIB_PATH=$V timeout 100 python tools/bench_k1.py 1000000; done
  echo default; timeout 120 python tools/bench_snapshots.py 1000000 64; echo aligned; SSB_LIB_PATH=$V timeout 120 python tools/bench_snapshots.py 1000000 64
  echo default; timeout 120 python tools/bench_snapshots.py 1000000 16; echo aligned; SSB_LIB_PATH=$V timeout 120 python tools/bench_snapshots.py 1000000 16
  echo default; timeout 100 python tools/bench_k1.py 1000000 8 c3; echo aligned; SSB_LIB_PATH=$V timeout 100 python tools/bench_k1.py 1000000 8 c3 ) > gpurun_out/o_align.log 2>&1
grep -v "^+" gpurun_out/o_align.log | cut -c1-200
