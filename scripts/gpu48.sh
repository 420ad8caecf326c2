cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2O_bench_n1.json 2> gpurun_out/r2O_bench_n1.err; tail -c 3000 gpurun_out/r2O_bench_n1.json; tail -2 gpurun_out/r2O_bench_n1.err
