# round-2 call R: bars confined to the generic kernels: same-box A/B of the hot kernels against the pre-bar build (build/variants/align_2cta.so:
# identical source for the final-state orbit kernel and the response kernel), then the bar tests
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$GRAFT_REPO_ROOT/build/variants/align_2cta.so
( for rep in 1 2; do echo "now"; timeout 100 python tools/bench_k1.py 1000000; echo "pre-bar"; SSB_LIB_PATH=$V timeout 100 python tools/bench_k1.py 1000000; done
  for cfg in "10000 1000 1e-6" "2000 1000 1e-11"; do echo "now $cfg"; timeout 100 python tools/bench_response.py $cfg; echo "pre-bar $cfg"; SSB_LIB_PATH=$V timeout 100 python tools/bench_response.py $cfg; done
  echo "now c3"; timeout 100 python tools/bench_k1.py 1000000 8 c3; echo "pre-bar c3"; SSB_LIB_PATH=$V timeout 100 python tools/bench_k1.py 1000000 8 c3
  echo "now snapshots"; timeout 120 python tools/bench_snapshots.py 1000000 64 ) > gpurun_out/r_regress.log 2>&1
grep -v "^+" gpurun_out/r_regress.log | cut -c1-170
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -W always -x -k "rotating_bars or closed_forms or third or variational or restricted" ) > gpurun_out/r_pytest_bars.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/r_pytest_bars.log | tail -3
grep -n "^E  " gpurun_out/r_pytest_bars.log | cut -c1-300 | head -20
