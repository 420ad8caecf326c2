cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NCB200_LIB=$GRAFT_REPO_ROOT/ncrystal_b200/lib/libncrystal_b200_tuning.so
for r in 1 2; do
for g in 1 0; do
  echo "== eval_groups=$g run $r" | tee -a gpurun_out/r2W_eval_ab.txt
  NCB200_SC_EVAL_GROUPS=$g timeout 300 python tests/ge_time.py 2>&1 | tail -3 | cut -c1-230 | tee -a gpurun_out/r2W_eval_ab.txt
done
done
unset NCB200_LIB
timeout 900 python -m pytest tests/test_gpu_parity_aniso.py tests/test_gpu_minimc.py tests/test_gpu_api_semantics.py -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_tables_and_handles.py -x -q -k "sweep or parity or alive" 2>&1 | tail -3
timeout 300 python tests/mmc_ktime.py Ge 1e6 2>&1 | tail -1 | cut -c1-330
