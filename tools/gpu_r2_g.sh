# round-2 call G: what one rank of an 8-way split stream costs on its own (no gather): 125 k / 250 k / 500 k-particle streams, pipeline on/off
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( for n in 125000 250000 500000 1000000; do for sp in 0 1; do echo "n=$n SSB_STREAM_SPLIT=$sp"; SSB_STREAM_SPLIT=$sp timeout 100 python tools/bench_k1.py $n; done; done ) > gpurun_out/g_shard.log 2>&1
grep -v "^+" gpurun_out/g_shard.log
( time timeout 900 python -m pytest tests -m gpu -q -W always ) > gpurun_out/g_pytest_gpu.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/g_pytest_gpu.log | tail
grep -n "^E  " gpurun_out/g_pytest_gpu.log | cut -c1-300 | head
