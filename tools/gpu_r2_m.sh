# round-2 call M: response kernel without phase barriers (named barriers, last-arriver controllers): bit-identity test, timings, ncu
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -W always -k "response or c4 or driver" ) > gpurun_out/m_pytest_resp.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/m_pytest_resp.log | tail
grep -n "^E  " gpurun_out/m_pytest_resp.log | cut -c1-300 | head -20
( for cfg in "10000 1000 1e-6" "2000 1000 1e-11" "100000 1000 1e-6"; do echo "$cfg"; timeout 200 python tools/bench_response.py $cfg; done
  for np in 4 8; do echo "SSB_RESP_NP=$np 10000 1000 1e-6"; SSB_RESP_NP=$np timeout 200 python tools/bench_response.py 10000 1000 1e-6; done
  echo "SSB_RESP_NP=8 production"; SSB_RESP_NP=8 timeout 200 python tools/bench_response.py 2000 1000 1e-11 ) > gpurun_out/m_response.log 2>&1
grep -v "^+" gpurun_out/m_response.log | cut -c1-150
timeout 300 ncu --set full --import-source on --clock-control none -k regex:response_kernel_mp -c 1 -f -o gpurun_out/m_resp python tools/bench_response.py 10000 1000 1e-6 > gpurun_out/m_ncu_resp.log 2>&1
timeout 100 python tools/ncu_summary.py gpurun_out/m_resp.ncu-rep > gpurun_out/m_resp_ncu.txt 2>&1
timeout 100 python tools/ncu_source_lines.py gpurun_out/m_resp.ncu-rep > gpurun_out/m_resp_source.txt 2>&1
grep -E "time_duration|dram__bytes|fp64_cycles|issue_active|stalled_(long|wait|no_inst|barrier|math|short)|inst_executed.sum|derived|warps_active" gpurun_out/m_resp_ncu.txt
head -30 gpurun_out/m_resp_source.txt
