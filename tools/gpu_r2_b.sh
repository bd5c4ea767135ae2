# round-2 validation call B: K3 with retired items (tests, A/B timing, ncu), snapshot-kernel A/B, bench
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -W always ) > gpurun_out/b_pytest_gpu.log 2>&1
tail -40 gpurun_out/b_pytest_gpu.log | cut -c1-700
( for r in 0 1; do echo "SSB_RESP_RETIRE=$r"; SSB_RESP_RETIRE=$r timeout 100 python tools/bench_response.py 10000 1000 1e-6; SSB_RESP_RETIRE=$r timeout 100 python tools/bench_response.py 2000 1000 1e-11; SSB_RESP_RETIRE=$r timeout 200 python tools/bench_response.py 100000 1000 1e-6; done ) > gpurun_out/b_response_ab.log 2>&1
cat gpurun_out/b_response_ab.log
( timeout 120 python tools/bench_snapshots.py 1000000 64; timeout 120 python tools/bench_snapshots.py 1000000 16; echo snap2; SSB_LIB_PATH=$GRAFT_REPO_ROOT/streamsculptor_b200/_lib/libssb200_snap2.so timeout 120 python tools/bench_snapshots.py 1000000 64; SSB_LIB_PATH=$GRAFT_REPO_ROOT/streamsculptor_b200/_lib/libssb200_snap2.so timeout 120 python tools/bench_snapshots.py 1000000 16 ) > gpurun_out/b_snapshots.log 2>&1
cat gpurun_out/b_snapshots.log
timeout 400 ncu --set full --import-source on --clock-control none -k regex:response_kernel_mp -s 1 -c 1 -f -o gpurun_out/b_resp python tools/bench_response.py 10000 1000 1e-6 > gpurun_out/b_ncu_resp.log 2>&1
timeout 100 python tools/ncu_summary.py gpurun_out/b_resp.ncu-rep > gpurun_out/b_resp_ncu.txt 2>&1
head -70 gpurun_out/b_resp_ncu.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:orbit_kernel -s 1 -c 1 -f -o gpurun_out/b_orbit_snap python tools/bench_snapshots.py 1000000 64 > gpurun_out/b_ncu_snap.log 2>&1
timeout 100 python tools/ncu_summary.py gpurun_out/b_orbit_snap.ncu-rep > gpurun_out/b_orbit_snap_ncu.txt 2>&1
timeout 400 python bench.py > gpurun_out/b_bench_n1.json 2> gpurun_out/b_bench_n1.err
cat gpurun_out/b_bench_n1.json | cut -c1-1500
ls -la gpurun_out
