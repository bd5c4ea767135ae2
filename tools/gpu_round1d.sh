set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
rm -f gpurun_out/ab.log
for rep in 1 2; do
for v in default k1old; do
  if [ $v = default ]; then unset SSB_LIB_PATH; else export SSB_LIB_PATH=$PWD/build/variants/$v.so; fi
  echo "== $v" >> gpurun_out/ab.log
  timeout 150 python tools/bench_k1.py >> gpurun_out/ab.log 2>&1
  timeout 150 python tools/bench_response.py 2000 1000 1e-11 >> gpurun_out/ab.log 2>&1
  timeout 150 python tools/bench_response.py 10000 1000 1e-6 >> gpurun_out/ab.log 2>&1
done
done
unset SSB_LIB_PATH
cat gpurun_out/ab.log
