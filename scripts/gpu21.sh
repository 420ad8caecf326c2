cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vdos.py -x -q 2>&1 | tail -5
timeout 900 python tests/vdos_time.py > gpurun_out/r2p_vdos_time.jsonl 2> gpurun_out/r2p_vdos_time.err; cat gpurun_out/r2p_vdos_time.jsonl; tail -5 gpurun_out/r2p_vdos_time.err
cat > /tmp/vd1.py <<'P'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import _vdos
from ncrystal_b200 import _lib
g = _vdos.load_golden()
p = _vdos.RawVdosAPI(_lib.lib())
s, m, _ = [float(x) for x in g["in_Al_meta"]]
p.kernel(g["in_Al_egrid"], g["in_Al_density"], s, m, 293.15, 3)
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_vdos_launches.csv python /tmp/vd1.py > gpurun_out/r2p_ncu_vdos.log 2>&1; tail -3 gpurun_out/r2p_ncu_vdos.log; wc -l gpurun_out/r2p_vdos_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_vdos -c 40 -o gpurun_out/r2p_vdos python /tmp/vd1.py > gpurun_out/r2p_ncu_vdos_full.log 2>&1; tail -3 gpurun_out/r2p_ncu_vdos_full.log
ncu -i gpurun_out/r2p_vdos.ncu-rep --page raw --csv > gpurun_out/r2p_vdos.raw.csv 2>/dev/null; wc -c gpurun_out/r2p_vdos.raw.csv
