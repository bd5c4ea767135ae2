# round-2 call U: ncu capture of the C3 orbit kernel (fused MW3 + one fast-extra moving Plummer), final build
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SSB_STREAM_SPLIT=0 timeout 300 ncu --set full --import-source on --clock-control none -k regex:orbit_kernel -s 2 -c 1 -f -o gpurun_out/u_c3 python tools/bench_k1.py 1000000 8 c3 > gpurun_out/u_ncu_c3.log 2>&1
timeout 100 python tools/ncu_summary.py gpurun_out/u_c3.ncu-rep > gpurun_out/u_c3_ncu.txt 2>&1
timeout 100 python tools/ncu_source_lines.py gpurun_out/u_c3.ncu-rep > gpurun_out/u_c3_source.txt 2>&1
rm -f gpurun_out/*.ncu-rep
grep -E "Kernel Name|time_duration|dram__bytes|fp64_cycles|issue_active|stalled_(long|wait|no_inst|math|short)|inst_executed.sum|derived|registers" gpurun_out/u_c3_ncu.txt | sed 's/smsp__average_warps_issue_//'
head -14 gpurun_out/u_c3_source.txt
