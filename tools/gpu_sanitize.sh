# compute-sanitizer memcheck + racecheck of the kernels changed in session 3 (tools/sanitize_new_kernels.py), then the response tests
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 280 compute-sanitizer --tool memcheck python tools/sanitize_new_kernels.py ) > gpurun_out/san_memcheck.log 2>&1
tail -2 gpurun_out/san_memcheck.log
( timeout 280 compute-sanitizer --tool racecheck python tools/sanitize_new_kernels.py ) > gpurun_out/san_racecheck.log 2>&1
tail -2 gpurun_out/san_racecheck.log
( timeout 300 python -m pytest tests -m gpu -x -q -k "response or host or restricted" ) > gpurun_out/pytest_gpu_resp.log 2>&1
tail -3 gpurun_out/pytest_gpu_resp.log
