cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tests/ge_time.py 2>&1 | tail -3 | tee gpurun_out/r2z_ge_time.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sc_find|k_sc_eval|k_sc_sample|k_classify_aniso" -c 8 -o gpurun_out/r2z_ge python tests/ge_time.py > gpurun_out/r2z_ncu.log 2>&1; tail -2 gpurun_out/r2z_ncu.log
ncu -i gpurun_out/r2z_ge.ncu-rep --page raw --csv > gpurun_out/r2z_ge.raw.csv 2>/dev/null; wc -c gpurun_out/r2z_ge.raw.csv
