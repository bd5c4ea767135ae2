# One gpurun call that validates a build on a B200 box: GPU parity tests, smoke(), bench.py, the ncu launch list of the bench, one
# full ncu capture of the orbit kernel and the per-config throughput tools.  Usage: gpurun --timeout 900 -- bash tools/gpu_validate.sh
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/b_ncu.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:orbit_kernel -s 4 -c 1 -f -o gpurun_out/orbit_k1 python bench.py --steps 2 --warmup 1 > gpurun_out/b_ncu2.log 2>&1
( timeout 300 python tools/bench_configs.py; timeout 100 python tools/bench_response.py 2000 1000 1e-11; timeout 100 python tools/bench_response.py 10000 1000 1e-6; timeout 100 python tools/bench_response.py 100000 1000 1e-6 ) > gpurun_out/configs.log 2>&1
cat gpurun_out/configs.log
# A/B of the two-part gen_stream pipeline on this box
( SSB_STREAM_SPLIT=0 timeout 100 python tools/bench_k1.py; SSB_STREAM_SPLIT=1 timeout 100 python tools/bench_k1.py ) > gpurun_out/split.log 2>&1
cat gpurun_out/split.log
ls -la gpurun_out
