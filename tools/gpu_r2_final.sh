# round-2 final validation: full GPU tests, smoke, bench N=1 (+ the reference arm), launch list, full ncu captures of the three hot kernels,
# per-config throughput, compute-sanitizer on the kernels of rounds 1-2
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/f_smi.txt 2>&1
( time timeout 1100 python -m pytest tests -m gpu -q -W always ) > gpurun_out/f_pytest_gpu.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/f_pytest_gpu.log | tail
grep -n "^E  " gpurun_out/f_pytest_gpu.log | cut -c1-300 | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; tail -2 gpurun_out/f_smoke.log
timeout 500 python bench.py > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err
cut -c1-700 gpurun_out/f_bench_n1.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
cut -c1-500 gpurun_out/f_bench_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/f_ncu_launches.log 2>&1
timeout 100 python tools/launch_shares.py gpurun_out/f_launches.csv "round 2: python bench.py --steps 2 --warmup 1 under ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: shares, not absolutes)" > gpurun_out/f_launch_shares.txt 2>&1
head -20 gpurun_out/f_launch_shares.txt
SSB_STREAM_SPLIT=0 timeout 300 ncu --set full --import-source on --clock-control none -k regex:orbit_kernel -s 2 -c 1 -f -o gpurun_out/f_orbit python tools/bench_k1.py 1000000 > gpurun_out/f_ncu_orbit.log 2>&1
timeout 100 python tools/ncu_summary.py gpurun_out/f_orbit.ncu-rep > gpurun_out/f_orbit_ncu.txt 2>&1
timeout 100 python tools/ncu_source_lines.py gpurun_out/f_orbit.ncu-rep > gpurun_out/f_orbit_source.txt 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:orbit_kernel -s 1 -c 1 -f -o gpurun_out/f_snap python tools/bench_snapshots.py 1000000 64 > gpurun_out/f_ncu_snap.log 2>&1
timeout 100 python tools/ncu_summary.py gpurun_out/f_snap.ncu-rep > gpurun_out/f_snap_ncu.txt 2>&1
timeout 100 python tools/ncu_source_lines.py gpurun_out/f_snap.ncu-rep > gpurun_out/f_snap_source.txt 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:response_kernel_mp -c 1 -f -o gpurun_out/f_resp python tools/bench_response.py 10000 1000 1e-6 > gpurun_out/f_ncu_resp.log 2>&1
timeout 100 python tools/ncu_summary.py gpurun_out/f_resp.ncu-rep > gpurun_out/f_resp_ncu.txt 2>&1
timeout 100 python tools/ncu_source_lines.py gpurun_out/f_resp.ncu-rep > gpurun_out/f_resp_source.txt 2>&1
for f in f_orbit f_snap f_resp; do echo $f; grep -E "Kernel Name|time_duration|dram__bytes|fp64_cycles|issue_active|stalled_(long|wait|no_inst|barrier|math|short)|inst_executed.sum|derived|warps_active" gpurun_out/${f}_ncu.txt | sed 's/smsp__average_warps_issue_//'; done
rm -f gpurun_out/*.ncu-rep          # summaries only travel back (64 MiB limit)
( timeout 300 python tools/bench_configs.py; timeout 100 python tools/bench_response.py 2000 1000 1e-11; timeout 100 python tools/bench_response.py 10000 1000 1e-6; timeout 100 python tools/bench_response.py 100000 1000 1e-6
  timeout 120 python tools/bench_snapshots.py 1000000 64; timeout 120 python tools/bench_snapshots.py 1000000 16; timeout 120 python tools/bench_snapshots.py 1000000 2 ) > gpurun_out/f_configs.log 2>&1
grep -v "^+" gpurun_out/f_configs.log | cut -c1-250
( timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_new_kernels.py ) > gpurun_out/f_san_memcheck.log 2>&1
tail -3 gpurun_out/f_san_memcheck.log
( timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_new_kernels.py ) > gpurun_out/f_san_racecheck.log 2>&1
grep -B2 -A12 -i "hazard\|Warning" gpurun_out/f_san_racecheck.log | head -60
tail -3 gpurun_out/f_san_racecheck.log
