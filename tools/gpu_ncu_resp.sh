set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 ncu --set full --import-source on --clock-control none -k regex:response_kernel -c 1 -f -o gpurun_out/resp_r1c python tools/bench_response.py 2000 1000 1e-11 > gpurun_out/ncu_resp.log 2>&1
tail -3 gpurun_out/ncu_resp.log
ls -la gpurun_out/
