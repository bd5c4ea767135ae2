cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
( timeout 250 python tools/bench_c5.py 1e7 1e5; timeout 100 python tools/bench_k1.py ) > gpurun_out/c5b.log 2>&1
cat gpurun_out/c5b.log
