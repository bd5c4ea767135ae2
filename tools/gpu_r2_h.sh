# round-2 call H: exclusive-SM progenitor solves in the stream pipeline: shard cost with / without
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( for n in 125000 250000 1000000; do for ex in 0 1; do echo "n=$n SSB_STREAM_EXCLUSIVE=$ex"; SSB_STREAM_EXCLUSIVE=$ex timeout 100 python tools/bench_k1.py $n; done; done ) > gpurun_out/h_shard.log 2>&1
grep -v "^+" gpurun_out/h_shard.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "stream" 2>&1 | tail -3
