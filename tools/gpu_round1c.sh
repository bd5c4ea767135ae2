set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
for v in default r64x6 r32x12 r96x4; do
  if [ $v = default ]; then unset SSB_LIB_PATH; else export SSB_LIB_PATH=$PWD/build/variants/$v.so; fi
  echo "== $v" >> gpurun_out/resp_variants.log
  timeout 150 python tools/bench_response.py 2000 1000 1e-11 >> gpurun_out/resp_variants.log 2>&1
  timeout 150 python tools/bench_response.py 10000 1000 1e-6 >> gpurun_out/resp_variants.log 2>&1
done
unset SSB_LIB_PATH
cat gpurun_out/resp_variants.log
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
