cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python tests/parity_sweep.py 5e6 2e6 > gpurun_out/r2B_parity_sweep.jsonl 2> gpurun_out/r2B_parity_sweep.err; cut -c1-260 gpurun_out/r2B_parity_sweep.jsonl; tail -2 gpurun_out/r2B_parity_sweep.err
