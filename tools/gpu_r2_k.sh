# round-2 call K: warp-autonomous response kernel: bit-identity test, A/B against the CTA-wide kernel, production-driver tests
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -W always -k "response or driver or impact or smoke" ) > gpurun_out/k_pytest_resp.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/k_pytest_resp.log | tail
grep -n "^E  " gpurun_out/k_pytest_resp.log | cut -c1-300 | head -20
( for k in mp wa; do for cfg in "10000 1000 1e-6" "2000 1000 1e-11" "100000 1000 1e-6"; do echo "SSB_RESP_KERNEL=$k $cfg"; SSB_RESP_KERNEL=$k timeout 200 python tools/bench_response.py $cfg; done; done
  echo "wa, 8 slots"; SSB_RESP_KERNEL=wa SSB_RESP_NP=8 timeout 200 python tools/bench_response.py 10000 1000 1e-6
  echo "wa, 4 slots"; SSB_RESP_KERNEL=wa SSB_RESP_NP=4 timeout 200 python tools/bench_response.py 10000 1000 1e-6
  echo "wa, 8 slots, production"; SSB_RESP_KERNEL=wa SSB_RESP_NP=8 timeout 200 python tools/bench_response.py 2000 1000 1e-11 ) > gpurun_out/k_response_ab.log 2>&1
grep -v "^+" gpurun_out/k_response_ab.log
SSB_RESP_KERNEL=wa timeout 300 ncu --set full --import-source on --clock-control none -k regex:response_kernel_wa -c 1 -f -o gpurun_out/k_resp_wa python tools/bench_response.py 10000 1000 1e-6 > gpurun_out/k_ncu_resp.log 2>&1
timeout 100 python tools/ncu_summary.py gpurun_out/k_resp_wa.ncu-rep > gpurun_out/k_resp_wa_ncu.txt 2>&1
timeout 100 python tools/ncu_source_lines.py gpurun_out/k_resp_wa.ncu-rep > gpurun_out/k_resp_wa_source.txt 2>&1
grep -E "time_duration|dram__bytes|fp64_cycles|issue_active|stalled_(long|wait|no_inst|barrier|math|short)|inst_executed.sum|derived|warps_active" gpurun_out/k_resp_wa_ncu.txt
head -30 gpurun_out/k_resp_wa_source.txt
