cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for r in 1 2; do
for v in old new; do
  if [ $v = old ]; then export NCB200_LIB=$GRAFT_REPO_ROOT/ncrystal_b200/lib/libncrystal_b200_old.so; else unset NCB200_LIB; fi
  echo "== $v run $r" | tee -a gpurun_out/r2S_ge_ab.txt
  timeout 300 python tests/ge_time.py 2>&1 | tail -3 | cut -c1-420 | tee -a gpurun_out/r2S_ge_ab.txt
done
done
for v in old new; do
  if [ $v = old ]; then export NCB200_LIB=$GRAFT_REPO_ROOT/ncrystal_b200/lib/libncrystal_b200_old.so; else unset NCB200_LIB; fi
  echo "== $v mmc" | tee -a gpurun_out/r2S_ge_ab.txt
  timeout 300 python tests/mmc_ktime.py Ge 1e6 2>&1 | tail -1 | cut -c1-330 | tee -a gpurun_out/r2S_ge_ab.txt
  timeout 300 python bench.py --config Ge --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | tail -1 | cut -c1-400 | tee -a gpurun_out/r2S_ge_ab.txt
done
unset NCB200_LIB; timeout 900 python -m pytest tests/test_gpu_parity_aniso.py tests/test_gpu_minimc.py -x -q 2>&1 | tail -3
