cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sc_find|k_sc_eval_groups|k_classify_aniso|k_sc_sample_threads" -s 12 -c 5 -o gpurun_out/r2T_ge python tests/ge_time.py > gpurun_out/r2T_ncu.log 2>&1; tail -2 gpurun_out/r2T_ncu.log
ls -la gpurun_out/r2T_ge.ncu-rep
