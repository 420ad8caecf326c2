cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_vdos.py tests/test_gpu_tables_and_handles.py -x -q 2>&1 | tail -12
