cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_vdos.py -x -q 2>&1 | tail -25
timeout 600 python tests/vdos_time.py > gpurun_out/r2p_vdos_time.jsonl 2> gpurun_out/r2p_vdos_time.err; cat gpurun_out/r2p_vdos_time.jsonl; tail -5 gpurun_out/r2p_vdos_time.err
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2p_pytest.log; cat gpurun_out/r2p_pytest.log
