cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python tests/vdos_sweep.py device 4 > gpurun_out/r2J_vdos_sweep_lux4.jsonl 2> gpurun_out/r2J_sweep.err; tail -1 gpurun_out/r2J_vdos_sweep_lux4.jsonl; tail -2 gpurun_out/r2J_sweep.err
timeout 2400 python tests/vdos_sweep.py device 5 6 > gpurun_out/r2J_vdos_sweep_lux5_every6.jsonl 2>> gpurun_out/r2J_sweep.err; tail -1 gpurun_out/r2J_vdos_sweep_lux5_every6.jsonl
