cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_minimc.py -x -q 2>&1 | tail -15
timeout 600 python tests/mmc_bench.py 1e7 1e6 > gpurun_out/r2o_mmc_bench.jsonl 2> gpurun_out/r2o_mmc_bench.err; cat gpurun_out/r2o_mmc_bench.jsonl | cut -c1-420; tail -3 gpurun_out/r2o_mmc_bench.err
