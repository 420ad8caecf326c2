cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tests/lc_time.py > gpurun_out/r2G_lc_time.json 2> gpurun_out/r2G_lc_time.err; cut -c150-520 gpurun_out/r2G_lc_time.json; tail -2 gpurun_out/r2G_lc_time.err
