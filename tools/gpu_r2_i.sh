# round-2 call I: tableau coefficients from the constant bank vs immediates: K1 A/B (stream, snapshots, response)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
IMM=$GRAFT_REPO_ROOT/streamsculptor_b200/_lib/libssb200_imm.so
( for rep in 1 2; do echo constbank; timeout 100 python tools/bench_k1.py 1000000; echo immediates; SSB_LIB_PATH=$IMM timeout 100 python tools/bench_k1.py 1000000; done
  echo constbank; timeout 100 python tools/bench_k1.py 1000000 5; echo immediates; SSB_LIB_PATH=$IMM timeout 100 python tools/bench_k1.py 1000000 5
  echo constbank; timeout 100 python tools/bench_k1.py 1000000 8 c3; echo immediates; SSB_LIB_PATH=$IMM timeout 100 python tools/bench_k1.py 1000000 8 c3
  echo constbank; timeout 120 python tools/bench_snapshots.py 1000000 64; echo immediates; SSB_LIB_PATH=$IMM timeout 120 python tools/bench_snapshots.py 1000000 64
  echo constbank; timeout 100 python tools/bench_response.py 10000 1000 1e-6; echo immediates; SSB_LIB_PATH=$IMM timeout 100 python tools/bench_response.py 10000 1000 1e-6 ) > gpurun_out/i_constbank.log 2>&1
grep -v "^+" gpurun_out/i_constbank.log
