cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 2400 python -m pytest tests -m gpu -x -q) > gpurun_out/r2k_pytest.log 2>&1; tail -25 gpurun_out/r2k_pytest.log
