# round-2 call L: out-of-line item force (instruction-cache footprint) A/B for both response kernels; production-driver test
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$GRAFT_REPO_ROOT/build/variants/itemcall.so
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -W always -k "driver or impact" ) > gpurun_out/l_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/l_pytest.log | tail
grep -n "^E  " gpurun_out/l_pytest.log | cut -c1-300 | head -20
( for k in mp wa; do for cfg in "10000 1000 1e-6" "2000 1000 1e-11"; do
    echo "inline SSB_RESP_KERNEL=$k $cfg"; SSB_RESP_KERNEL=$k timeout 200 python tools/bench_response.py $cfg
    echo "itemcall SSB_RESP_KERNEL=$k $cfg"; SSB_LIB_PATH=$V SSB_RESP_KERNEL=$k timeout 200 python tools/bench_response.py $cfg; done; done
  echo "itemcall mp 1e5"; SSB_LIB_PATH=$V SSB_RESP_KERNEL=mp timeout 200 python tools/bench_response.py 100000 1000 1e-6
  echo "itemcall wa 1e5"; SSB_LIB_PATH=$V SSB_RESP_KERNEL=wa timeout 200 python tools/bench_response.py 100000 1000 1e-6 ) > gpurun_out/l_response_ab.log 2>&1
grep -v "^+" gpurun_out/l_response_ab.log | cut -c1-150
