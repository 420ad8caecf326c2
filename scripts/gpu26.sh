cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vdos.py -x -q 2>&1 | tail -4
timeout 1200 python tests/vdos_sweep.py device 1 > gpurun_out/r2u_vdos_sweep_lux1.jsonl 2> gpurun_out/r2u_sweep.err; tail -1 gpurun_out/r2u_vdos_sweep_lux1.jsonl; tail -2 gpurun_out/r2u_sweep.err
timeout 1800 python tests/vdos_sweep.py device 3 > gpurun_out/r2u_vdos_sweep_lux3.jsonl 2>> gpurun_out/r2u_sweep.err; tail -1 gpurun_out/r2u_vdos_sweep_lux3.jsonl
