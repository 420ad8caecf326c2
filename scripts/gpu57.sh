cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 700 python tests/parity_sweep.py 1e6 5e6 Ge,Cu_sc,Ge_vapi,PG > gpurun_out/r2Z_parity_oriented.jsonl 2> gpurun_out/r2Z_parity_oriented.err; cut -c1-400 gpurun_out/r2Z_parity_oriented.jsonl; tail -2 gpurun_out/r2Z_parity_oriented.err
