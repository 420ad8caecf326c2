cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_aniso.py tests/test_gpu_parity_extra.py -x -q 2>&1 | tail -4
timeout 600 python tests/lc_time.py > gpurun_out/r2E_lc_time.json 2> gpurun_out/r2E_lc_time.err; cut -c1-700 gpurun_out/r2E_lc_time.json; tail -2 gpurun_out/r2E_lc_time.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
