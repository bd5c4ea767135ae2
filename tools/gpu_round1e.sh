set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
rm -f gpurun_out/np.log
for np in 0 1 2 4; do
  echo "== SSB_RESP_NP=$np" >> gpurun_out/np.log
  SSB_RESP_NP=$np timeout 150 python tools/bench_response.py 2000 1000 1e-11 >> gpurun_out/np.log 2>&1
  SSB_RESP_NP=$np timeout 150 python tools/bench_response.py 10000 1000 1e-6 >> gpurun_out/np.log 2>&1
done
cat gpurun_out/np.log
