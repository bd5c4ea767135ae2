# round-2: ncu capture of the final response kernel
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:response_kernel_mp -c 1 -f -o gpurun_out/k3p python tools/bench_response.py 10000 1000 1e-6 > gpurun_out/k3p_ncu.log 2>&1
timeout 100 python tools/ncu_summary.py gpurun_out/k3p.ncu-rep > gpurun_out/k3p_ncu.txt 2>&1
timeout 100 python tools/ncu_source_lines.py gpurun_out/k3p.ncu-rep > gpurun_out/k3p_source.txt 2>&1
rm -f gpurun_out/*.ncu-rep
grep -E "Kernel Name|time_duration|dram__bytes|fp64_cycles|issue_active|stalled_(long|wait|no_inst|barrier|math|short|branch)|inst_executed.sum|derived|warps_active|local_op_ld.sum" gpurun_out/k3p_ncu.txt | sed 's/smsp__average_warps_issue_//'
head -14 gpurun_out/k3p_source.txt
