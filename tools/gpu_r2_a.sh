# round-2 validation call A: GPU tests, smoke, bench (N=1), snapshot path timing + ncu captures.
# Usage: gpurun --timeout 1500 -- bash tools/gpu_r2_a.sh
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt
( time timeout 900 python -m pytest tests -m gpu -q -W always ) > gpurun_out/a_pytest_gpu.log 2>&1
tail -60 gpurun_out/a_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; tail -2 gpurun_out/a_smoke.log
timeout 400 python bench.py > gpurun_out/a_bench_n1.json 2> gpurun_out/a_bench_n1.err
cat gpurun_out/a_bench_n1.json; tail -5 gpurun_out/a_bench_n1.err
( timeout 120 python tools/bench_snapshots.py 1000000 64; timeout 120 python tools/bench_snapshots.py 1000000 16 ) > gpurun_out/a_snapshots.log 2>&1
cat gpurun_out/a_snapshots.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:orbit_kernel -s 1 -c 1 -f -o gpurun_out/a_orbit_snap python tools/bench_snapshots.py 1000000 64 > gpurun_out/a_ncu_snap.log 2>&1
timeout 100 python tools/ncu_summary.py gpurun_out/a_orbit_snap.ncu-rep > gpurun_out/a_orbit_snap_ncu.txt 2>&1
cat gpurun_out/a_orbit_snap_ncu.txt | head -60
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/a_b_ncu.log 2>&1
ls -la gpurun_out
