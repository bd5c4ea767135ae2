set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
SSB_RESP_NP=4 timeout 400 ncu --set full --import-source on --clock-control none -k regex:response_kernel_mp -c 1 -f -o gpurun_out/resp_mp4 python tools/bench_response.py 10000 1000 1e-6 > gpurun_out/ncu_resp_mp.log 2>&1
tail -3 gpurun_out/ncu_resp_mp.log
