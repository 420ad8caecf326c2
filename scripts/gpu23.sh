cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for m in 1 2; do timeout 600 python tests/mmc_ktime.py Ge 1e6 $m 2>&1 | tail -1 | cut -c1-500 | tee -a gpurun_out/r2r_mmc_ktime.jsonl; done
for m in 1 2; do timeout 600 python tests/mmc_ktime.py Al 1e7 $m 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/r2r_mmc_ktime.jsonl; done
timeout 900 python -m pytest tests/test_gpu_minimc.py -x -q 2>&1 | tail -5
