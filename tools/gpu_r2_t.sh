# round-2 call T: tableau operands (constant bank vs immediates) re-measured with the unsplit build: cb1 = constant bank in every TU (tests the
# orbit kernels), cb0 = immediates in every TU (tests the response / saving kernels); main = immediates in ssb_kernels.cu, constant bank elsewhere
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B=$GRAFT_REPO_ROOT/build/variants
( for rep in 1 2; do echo "main"; timeout 100 python tools/bench_k1.py 1000000; echo "cb1"; SSB_LIB_PATH=$B/cb1.so timeout 100 python tools/bench_k1.py 1000000; done
  echo "main c3"; timeout 100 python tools/bench_k1.py 1000000 8 c3; echo "cb1 c3"; SSB_LIB_PATH=$B/cb1.so timeout 100 python tools/bench_k1.py 1000000 8 c3
  echo "main dopri5"; timeout 100 python tools/bench_k1.py 1000000 5; echo "cb1 dopri5"; SSB_LIB_PATH=$B/cb1.so timeout 100 python tools/bench_k1.py 1000000 5
  echo "main snapshots"; timeout 120 python tools/bench_snapshots.py 1000000 64; echo "cb1 snapshots"; SSB_LIB_PATH=$B/cb1.so timeout 120 python tools/bench_snapshots.py 1000000 64
  for cfg in "10000 1000 1e-6" "2000 1000 1e-11"; do echo "main $cfg"; timeout 100 python tools/bench_response.py $cfg; echo "cb0 $cfg"; SSB_LIB_PATH=$B/cb0.so timeout 100 python tools/bench_response.py $cfg; done ) > gpurun_out/t_constbank.log 2>&1
grep -v "^+" gpurun_out/t_constbank.log | cut -c1-150
