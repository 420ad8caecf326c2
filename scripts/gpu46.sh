cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2M_pytest.log; cat gpurun_out/r2M_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1300 python tests/vdos_sweep.py device 5 > gpurun_out/r2M_vdos_sweep_lux5_all.jsonl 2> gpurun_out/r2M_sweep.err; tail -1 gpurun_out/r2M_vdos_sweep_lux5_all.jsonl; tail -2 gpurun_out/r2M_sweep.err
