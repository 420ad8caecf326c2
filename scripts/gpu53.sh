cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2X_pytest.log; cat gpurun_out/r2X_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2X_bench_n1.json 2> gpurun_out/r2X_bench_n1.err; cut -c1-300 gpurun_out/r2X_bench_n1.json; tail -2 gpurun_out/r2X_bench_n1.err
