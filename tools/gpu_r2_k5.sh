# round-2: ncu capture of the shared-step attempt kernel (K5), 1e7 tracers, restricted N-body field; and with the 100-perturber set (variant PS)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:shared_attempt -s 10 -c 1 -f -o gpurun_out/k5 python tools/bench_c5.py 10000000 1000 > gpurun_out/k5_ncu.log 2>&1
timeout 100 python tools/ncu_summary.py gpurun_out/k5.ncu-rep > gpurun_out/k5_ncu.txt 2>&1
timeout 100 python tools/ncu_source_lines.py gpurun_out/k5.ncu-rep > gpurun_out/k5_source.txt 2>&1
rm -f gpurun_out/*.ncu-rep
grep -E "Kernel Name|time_duration|dram__bytes|fp64_cycles|issue_active|stalled_(long|wait|no_inst|math|short|barrier)|inst_executed.sum|derived|registers" gpurun_out/k5_ncu.txt | sed 's/smsp__average_warps_issue_//'
head -12 gpurun_out/k5_source.txt
