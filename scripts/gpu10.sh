cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests/test_gpu_parity_extra.py tests/test_gpu_parity_aniso.py -x -q -k "oriented or layered or aniso") > gpurun_out/r2j_pytest_lc.log 2>&1; tail -15 gpurun_out/r2j_pytest_lc.log
timeout 600 python tests/lc_time.py 2000000 > gpurun_out/r2j_lc_time.json 2> gpurun_out/r2j_lc_time.err; cat gpurun_out/r2j_lc_time.json; tail -3 gpurun_out/r2j_lc_time.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lc_scan -s 1 -c 1 -o gpurun_out/r2j_lc_scan python tests/lc_time.py 500000 > gpurun_out/r2j_ncu_lc.log 2>&1
ncu -i gpurun_out/r2j_lc_scan.ncu-rep --page raw --csv > gpurun_out/r2j_lc_scan.raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
