# round-2 call V: LMCPotential / not-a-knot track test, then the whole GPU suite once more
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -W always -x -k "lmc_potential" ) > gpurun_out/v_pytest_lmc.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/v_pytest_lmc.log | tail -3
grep -n "^E  " gpurun_out/v_pytest_lmc.log | cut -c1-300 | head -20
( time timeout 1100 python -m pytest tests -m gpu -q -W always ) > gpurun_out/v_pytest_gpu.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/v_pytest_gpu.log | tail
