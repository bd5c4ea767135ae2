# round-2 call Z3: why the response kernel is slow with the production base potential: retire on/off, one-particle kernel, softer progenitor, linear track
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( echo "plain"; timeout 100 python tools/bench_response.py 2000 1000 1e-11
  echo "prog"; timeout 100 python tools/bench_response.py 2000 1000 1e-11 prog
  echo "prog RETIRE=0"; SSB_RESP_RETIRE=0 timeout 100 python tools/bench_response.py 2000 1000 1e-11 prog
  echo "plain RETIRE=0"; SSB_RESP_RETIRE=0 timeout 100 python tools/bench_response.py 2000 1000 1e-11
  echo "prog NP=0"; SSB_RESP_NP=0 timeout 100 python tools/bench_response.py 2000 1000 1e-11 prog
  echo "prog r_s=0.1"; timeout 100 python tools/bench_response.py 2000 1000 1e-11 prog 0.1
  echo "prog linear"; timeout 100 python tools/bench_response.py 2000 1000 1e-11 prog 0.004 linear ) > gpurun_out/z3.log 2>&1
grep -v "^+" gpurun_out/z3.log | cut -c1-200
