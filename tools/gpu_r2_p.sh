# round-2 call P: saving kernel with the CTA-aligned loop: inlined force / out-of-line force / 2 CTAs per SM
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( for v in default align_ool align_2cta; do for m in 64 16; do echo "$v M=$m"; if [ $v = default ]; then timeout 120 python tools/bench_snapshots.py 1000000 $m; else SSB_LIB_PATH=$GRAFT_REPO_ROOT/build/variants/$v.so timeout 120 python tools/bench_snapshots.py 1000000 $m; fi; done; done ) > gpurun_out/p_snap.log 2>&1
grep -v "^+" gpurun_out/p_snap.log | cut -c1-200
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -W always -x -k "snapshots or dense or failure or goldens" ) > gpurun_out/p_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/p_pytest.log | tail -3
