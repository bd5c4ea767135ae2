cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 250 python tools/bench_c5.py 1e6 1e6; timeout 250 python tools/bench_c5.py 1e7 1e5 ) > gpurun_out/c5.log 2>&1
cat gpurun_out/c5.log
