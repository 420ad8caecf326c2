cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tests/mmc_ktime.py Ge 1e7 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r2r_mmc_ktime.jsonl
timeout 600 python tests/mmc_bench.py 1e7 1e6 > gpurun_out/r2s_mmc_bench.jsonl 2> gpurun_out/r2s_mmc_bench.err; cut -c1-330 gpurun_out/r2s_mmc_bench.jsonl
( echo "## memcheck (tests/sanitizer_vdos.py all)"; timeout 1200 compute-sanitizer --tool memcheck python tests/sanitizer_vdos.py all 2>&1 | grep -v "^$" | tail -12
  echo "## racecheck (tests/sanitizer_vdos.py all)"; timeout 1500 compute-sanitizer --tool racecheck python tests/sanitizer_vdos.py all 2>&1 | grep -v "^$" | tail -12
  echo "## initcheck (tests/sanitizer_vdos.py)"; timeout 900 compute-sanitizer --tool initcheck python tests/sanitizer_vdos.py 2>&1 | grep -v "^$" | tail -8 ) > gpurun_out/r2s_sanitizer.txt 2>&1
cat gpurun_out/r2s_sanitizer.txt
