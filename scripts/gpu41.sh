cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python tests/parity_sweep.py 5e7 1e7 Al,CH2,H2O,YAG,Ge > gpurun_out/r2H_parity_sweep_5e7.jsonl 2> gpurun_out/r2H_parity_sweep.err; cut -c1-260 gpurun_out/r2H_parity_sweep_5e7.jsonl; tail -2 gpurun_out/r2H_parity_sweep.err
