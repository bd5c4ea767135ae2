# round-2 call E: full GPU tests with the 4-part stream pipeline, snapshot kernel with staged coefficients, bench N=1
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -W always ) > gpurun_out/e_pytest_gpu.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/e_pytest_gpu.log | tail
( timeout 120 python tools/bench_snapshots.py 1000000 64; timeout 120 python tools/bench_snapshots.py 1000000 16; timeout 120 python tools/bench_snapshots.py 1000000 2 ) > gpurun_out/e_snapshots.log 2>&1
grep -v "^+" gpurun_out/e_snapshots.log
( SSB_STREAM_SPLIT=0 timeout 100 python tools/bench_k1.py; SSB_STREAM_SPLIT=1 timeout 100 python tools/bench_k1.py ) > gpurun_out/e_split.log 2>&1
grep -v "^+" gpurun_out/e_split.log
timeout 400 python bench.py > gpurun_out/e_bench_n1.json 2> gpurun_out/e_bench_n1.err
cut -c1-700 gpurun_out/e_bench_n1.json
timeout 300 ncu --set full --import-source on --clock-control none -k regex:orbit_kernel -s 1 -c 1 -f -o gpurun_out/e_orbit_snap python tools/bench_snapshots.py 1000000 64 > gpurun_out/e_ncu_snap.log 2>&1
timeout 100 python tools/ncu_summary.py gpurun_out/e_orbit_snap.ncu-rep > gpurun_out/e_orbit_snap_ncu.txt 2>&1
grep -E "time_duration|dram__bytes|fp64_cycles|issue_active|stalled_(long|wait|no_inst|math|short)|inst_executed.sum|derived|local_op_ld.sum" gpurun_out/e_orbit_snap_ncu.txt
