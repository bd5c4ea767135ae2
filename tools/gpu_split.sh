cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
( SSB_STREAM_SPLIT=0 timeout 100 python tools/bench_k1.py; SSB_STREAM_SPLIT=1 timeout 100 python tools/bench_k1.py; SSB_STREAM_SPLIT=0 timeout 100 python tools/bench_k1.py; SSB_STREAM_SPLIT=1 timeout 100 python tools/bench_k1.py ) > gpurun_out/split.log 2>&1
cat gpurun_out/split.log
timeout 200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json | cut -c1-300; tail -2 gpurun_out/bench_n1.err
