# round-2: C3 with the perturber on a cubic track (interpreter path) vs a linear one (inline fast extra)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 100 python tools/bench_k1.py 1000000 8 c3; timeout 200 python tools/bench_k1.py 1000000 8 c3cubic ) > gpurun_out/c3c.log 2>&1
grep -v "^+" gpurun_out/c3c.log | cut -c1-150
