# round-2 call C: K3 with up to 16 slots per CTA: response tests, slot-count A/B, ncu of the default
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_adaptive_parity.py tests/test_gpu_fullsize.py -m gpu -q -W always -k "response or lock_step or c4 or second_order or perturbation" ) > gpurun_out/c_pytest_resp.log 2>&1
tail -15 gpurun_out/c_pytest_resp.log | cut -c1-400
( for np in 4 8 16; do echo "SSB_RESP_NP=$np"; SSB_RESP_NP=$np timeout 100 python tools/bench_response.py 10000 1000 1e-6; SSB_RESP_NP=$np timeout 100 python tools/bench_response.py 2000 1000 1e-11; SSB_RESP_NP=$np timeout 200 python tools/bench_response.py 100000 1000 1e-6; done; echo default; timeout 100 python tools/bench_response.py 10000 1000 1e-6 ) > gpurun_out/c_response_np.log 2>&1
grep -v "^+" gpurun_out/c_response_np.log
timeout 400 ncu --set full --import-source on --clock-control none -k regex:response_kernel_mp -s 1 -c 1 -f -o gpurun_out/c_resp python tools/bench_response.py 10000 1000 1e-6 > gpurun_out/c_ncu_resp.log 2>&1
timeout 100 python tools/ncu_summary.py gpurun_out/c_resp.ncu-rep > gpurun_out/c_resp_ncu.txt 2>&1
grep -E "time_duration|dram__bytes|fp64_cycles|issue_active|stalled_(long|wait|no_inst|barrier|math|short)|inst_executed.sum|derived" gpurun_out/c_resp_ncu.txt
timeout 200 python tools/ncu_source_lines.py gpurun_out/c_resp.ncu-rep > gpurun_out/c_resp_source.txt 2>&1
head -50 gpurun_out/c_resp_source.txt
ls -la gpurun_out | tail -8
