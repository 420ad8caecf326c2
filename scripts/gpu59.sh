cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tests/mmc_ktime.py Ge 1e6 2>&1 | tail -1 | cut -c1-330
timeout 300 python tests/mmc_ktime.py Ge 1e6 2>&1 | tail -1 | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_minimc.py tests/test_gpu_parity_aniso.py -x -q 2>&1 | tail -2
