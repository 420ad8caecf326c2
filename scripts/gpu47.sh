cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vdos.py -x -q 2>&1 | tail -3
timeout 900 python tests/vdos_time.py > gpurun_out/r2N_vdos_time.jsonl 2> gpurun_out/r2N_vdos_time.err; cut -c1-230 gpurun_out/r2N_vdos_time.jsonl; tail -2 gpurun_out/r2N_vdos_time.err
timeout 1200 python tests/vdos_sweep.py device 3 > gpurun_out/r2N_vdos_sweep_lux3.jsonl 2> gpurun_out/r2N_sweep.err; tail -1 gpurun_out/r2N_vdos_sweep_lux3.jsonl
