cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( echo "## memcheck tests/sanitizer_ge.py (single-crystal kernels of round 2: packed-record search, eight-lane evaluation, thread-per-neutron sampling)"; timeout 900 compute-sanitizer --tool memcheck python tests/sanitizer_ge.py 2>&1 | grep -v "^$" | tail -6
  echo "## racecheck tests/sanitizer_ge.py"; timeout 1200 compute-sanitizer --tool racecheck python tests/sanitizer_ge.py 2>&1 | grep -v "^$" | tail -6
  echo "## memcheck tests/sanitizer_vdos.py all"; timeout 900 compute-sanitizer --tool memcheck python tests/sanitizer_vdos.py all 2>&1 | grep -v "^$" | tail -4 ) > gpurun_out/r2C_sanitizer.txt 2>&1
cat gpurun_out/r2C_sanitizer.txt
