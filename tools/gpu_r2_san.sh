# compute-sanitizer memcheck + racecheck on the final build (tools/sanitize_new_kernels.py)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_new_kernels.py ) > gpurun_out/san_memcheck.log 2>&1
tail -4 gpurun_out/san_memcheck.log
( timeout 700 compute-sanitizer --tool racecheck python tools/sanitize_new_kernels.py ) > gpurun_out/san_racecheck.log 2>&1
grep -B2 -A8 -i "hazard\|Warning" gpurun_out/san_racecheck.log | head -40
tail -3 gpurun_out/san_racecheck.log
