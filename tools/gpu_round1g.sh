set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q -k "response or host" ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
rm -f gpurun_out/l2.log
for v in default nohint pf2 nohint_pf2; do
  if [ $v = default ]; then unset SSB_LIB_PATH; else export SSB_LIB_PATH=$PWD/build/variants/$v.so; fi
  for np in 4 2 0; do
    echo "== $v NP=$np" >> gpurun_out/l2.log
    SSB_RESP_NP=$np timeout 150 python tools/bench_response.py 2000 1000 1e-11 >> gpurun_out/l2.log 2>&1
    SSB_RESP_NP=$np timeout 150 python tools/bench_response.py 10000 1000 1e-6 >> gpurun_out/l2.log 2>&1
  done
done
unset SSB_LIB_PATH
for np in 4 0; do
  echo "== default NP=$np 100000" >> gpurun_out/l2.log
  SSB_RESP_NP=$np timeout 200 python tools/bench_response.py 100000 1000 1e-6 >> gpurun_out/l2.log 2>&1
done
cat gpurun_out/l2.log
