cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2Z_bench_n2.json 2> gpurun_out/r2Z_bench_n2.err; tail -1 gpurun_out/r2Z_bench_n2.json | cut -c1-260; tail -2 gpurun_out/r2Z_bench_n2.err | cut -c1-300
