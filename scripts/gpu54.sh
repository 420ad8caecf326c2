cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for r in 1 2; do timeout 300 python tests/ge_time.py 2>&1 | tail -3 | cut -c1-200 | tee -a gpurun_out/r2Z2_ge_time.txt; done
timeout 300 python tests/mmc_ktime.py Ge 1e6 2>&1 | tail -1 | cut -c1-330
timeout 900 python -m pytest tests/test_gpu_parity_aniso.py tests/test_gpu_minimc.py tests/test_gpu_api_semantics.py -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_tables_and_handles.py -x -q -k "sweep or parity" 2>&1 | tail -3
