cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_minimc.py tests/test_gpu_parity_aniso.py -x -q 2>&1 | tail -3
for a in "Ge 1e6" "Ge 1e7" "Al 1e7"; do timeout 600 python tests/mmc_ktime.py $a 2>&1 | tail -1 | cut -c1-330 | tee -a gpurun_out/r2w_mmc_ktime.jsonl; done
python - <<'P'
import sys, numpy as np
sys.path.insert(0,'tests')
import _vdos
g=_vdos.load_golden()
for c in ("Al","CH2_H","Be"):
    np.concatenate([g['in_%s_egrid'%c],g['in_%s_meta'%c],g['in_%s_density'%c]]).tofile('/tmp/%s.bin'%c)
P
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -DNCB_VDOS_TIMING -Incrystal_b200/csrc tests/tools/vdos_stage_times.cu ncrystal_b200/csrc/ncb_vdos.cu -o /tmp/vdos_stage_times -ccbin /usr/bin/g++ 2>&1 | tail -2
/tmp/vdos_stage_times /tmp/Al.bin 3 2>&1 | tail -22 | tee gpurun_out/r2w_vdos_stage_times.txt
/tmp/vdos_stage_times /tmp/Be.bin 4 2>&1 | tail -11 | tee -a gpurun_out/r2w_vdos_stage_times.txt
