cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_mmc_tail -c 1 -o gpurun_out/r2t_tail python tests/mmc_one.py Ge 1e6 1 3.2 > gpurun_out/r2t_ncu.log 2>&1; tail -3 gpurun_out/r2t_ncu.log
ncu -i gpurun_out/r2t_tail.ncu-rep --page raw --csv > gpurun_out/r2t_tail.raw.csv 2>/dev/null
ncu -i gpurun_out/r2t_tail.ncu-rep --page source --csv > gpurun_out/r2t_tail.source.csv 2>/dev/null; wc -c gpurun_out/r2t_tail.source.csv
