# round-2 call Y: response kernel with the production driver's base potential (galaxy + moving Plummer progenitor on a cubic track)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 100 python tools/bench_response.py 10000 1000 1e-6; timeout 100 python tools/bench_response.py 10000 1000 1e-6 prog
  timeout 100 python tools/bench_response.py 2000 1000 1e-11; timeout 100 python tools/bench_response.py 2000 1000 1e-11 prog ) > gpurun_out/y_resp_prog.log 2>&1
grep -v "^+" gpurun_out/y_resp_prog.log | cut -c1-200
