# round-2 call D: full GPU tests, snapshot-kernel A/B (out-of-line vs inlined force), response defaults, bench N=1
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -W always ) > gpurun_out/d_pytest_gpu.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/d_pytest_gpu.log | tail
( timeout 120 python tools/bench_snapshots.py 1000000 64; timeout 120 python tools/bench_snapshots.py 1000000 16; timeout 120 python tools/bench_snapshots.py 1000000 2; echo snapinl; SSB_LIB_PATH=$GRAFT_REPO_ROOT/streamsculptor_b200/_lib/libssb200_snapinl.so timeout 120 python tools/bench_snapshots.py 1000000 64; SSB_LIB_PATH=$GRAFT_REPO_ROOT/streamsculptor_b200/_lib/libssb200_snapinl.so timeout 120 python tools/bench_snapshots.py 1000000 16 ) > gpurun_out/d_snapshots.log 2>&1
grep -v "^+" gpurun_out/d_snapshots.log
( timeout 100 python tools/bench_response.py 10000 1000 1e-6; timeout 100 python tools/bench_response.py 2000 1000 1e-11; timeout 200 python tools/bench_response.py 100000 1000 1e-6 ) > gpurun_out/d_response.log 2>&1
grep -v "^+" gpurun_out/d_response.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:orbit_kernel -s 1 -c 1 -f -o gpurun_out/d_orbit_snap python tools/bench_snapshots.py 1000000 64 > gpurun_out/d_ncu_snap.log 2>&1
timeout 100 python tools/ncu_summary.py gpurun_out/d_orbit_snap.ncu-rep > gpurun_out/d_orbit_snap_ncu.txt 2>&1
grep -E "time_duration|dram__bytes|fp64_cycles|issue_active|stalled_(long|wait|no_inst|barrier|math|short)|inst_executed.sum|derived|local_op_ld.sum" gpurun_out/d_orbit_snap_ncu.txt
timeout 400 python bench.py > gpurun_out/d_bench_n1.json 2> gpurun_out/d_bench_n1.err
cut -c1-600 gpurun_out/d_bench_n1.json
