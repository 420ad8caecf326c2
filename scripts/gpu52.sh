cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sc_find|k_sc_eval_flat|k_classify_aniso|k_sc_sample_threads" -s 12 -c 5 -o gpurun_out/r2Y_ge python tests/ge_time.py > gpurun_out/r2Y_ncu.log 2>&1; tail -2 gpurun_out/r2Y_ncu.log
ls -la gpurun_out/r2Y_ge.ncu-rep
