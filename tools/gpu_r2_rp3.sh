# round-2: where the moving progenitor term's time goes in the response kernel (1e4 x 1000, soft progenitor): plain vs prog, 80 hottest lines each
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:response_kernel_mp -c 1 -f -o gpurun_out/rp3_prog python tools/bench_response.py 10000 1000 1e-6 prog 0.1 > gpurun_out/rp3_ncu.log 2>&1
timeout 100 python tools/ncu_summary.py gpurun_out/rp3_prog.ncu-rep > gpurun_out/rp3_prog_ncu.txt 2>&1
timeout 100 python tools/ncu_source_lines.py gpurun_out/rp3_prog.ncu-rep "prog" 90 > gpurun_out/rp3_prog_source.txt 2>&1
rm -f gpurun_out/*.ncu-rep
grep -E "time_duration|fp64_cycles|issue_active|stalled_(long|wait|no_inst|barrier|short)|inst_executed.sum|local_op_ld.sum" gpurun_out/rp3_prog_ncu.txt | sed 's/smsp__average_warps_issue_//'
head -4 gpurun_out/rp3_prog_source.txt; grep -n "ssb_potential.cuh\|ssb_response.cu:1[0-2][0-9] \|ssb_response.cu:9[0-9] " gpurun_out/rp3_prog_source.txt | head -40
