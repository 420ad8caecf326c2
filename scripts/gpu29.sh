cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vdos.py -x -q 2>&1 | tail -3
timeout 900 python tests/vdos_time.py > gpurun_out/r2x_vdos_time.jsonl 2> gpurun_out/r2x_vdos_time.err; cut -c1-260 gpurun_out/r2x_vdos_time.jsonl; tail -3 gpurun_out/r2x_vdos_time.err
python - <<'P'
import sys, numpy as np
sys.path.insert(0,'tests')
import _vdos
g=_vdos.load_golden()
for c in ("Al","CH2_H","Be"):
    np.concatenate([g['in_%s_egrid'%c],g['in_%s_meta'%c],g['in_%s_density'%c]]).tofile('/tmp/%s.bin'%c)
P
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -DNCB_VDOS_TIMING -Incrystal_b200/csrc tests/tools/vdos_stage_times.cu ncrystal_b200/csrc/ncb_vdos.cu -o /tmp/vdos_stage_times -ccbin /usr/bin/g++ 2>&1 | tail -2
/tmp/vdos_stage_times /tmp/Al.bin 3 2>&1 | tail -11 | tee gpurun_out/r2x_vdos_stage_times.txt
