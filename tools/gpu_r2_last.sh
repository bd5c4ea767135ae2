# last check of the round: the whole GPU suite and smoke() on the library in the tree
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -x ) > gpurun_out/last_pytest.log 2>&1
tail -2 gpurun_out/last_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last_smoke.log 2>&1; tail -1 gpurun_out/last_smoke.log
