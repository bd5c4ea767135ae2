# round-2: response kernel with 254 registers / 2 CTAs per SM (no spills) vs 168 registers / 3 CTAs
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$GRAFT_REPO_ROOT/build/variants/k3_2cta.so
( for cfg in "10000 1000 1e-6" "2000 1000 1e-11" "100000 1000 1e-6"; do echo "3 CTAs/SM: $cfg"; timeout 100 python tools/bench_response.py $cfg; echo "2 CTAs/SM: $cfg"; SSB_LIB_PATH=$V timeout 100 python tools/bench_response.py $cfg; done ) > gpurun_out/k32.log 2>&1
grep -v "^+" gpurun_out/k32.log | grep "CTAs\|^C4" | cut -c1-120
