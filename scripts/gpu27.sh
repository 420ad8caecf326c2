cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_minimc.py tests/test_gpu_parity_aniso.py tests/test_gpu_vdos.py -x -q 2>&1 | tail -5
for a in "Ge 1e6" "Ge 1e7" "Al 1e7"; do timeout 600 python tests/mmc_ktime.py $a 2>&1 | tail -1 | cut -c1-420 | tee -a gpurun_out/r2v_mmc_ktime.jsonl; done
