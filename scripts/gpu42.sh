cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NCB200_LIB=$PWD/ncrystal_b200/libv/exp_loglin.so
timeout 900 python -m pytest tests/test_gpu_parity_iso.py -x -q 2>&1 | tail -4
python bench.py --no-cpu-baseline --no-other-configs > gpurun_out/r2I_bench_loglin.json 2>/dev/null
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2I_bench_loglin.json').read().strip().splitlines()[-1])
print('value %.3e'%d['value'],'ms/step %.4f'%d['ms_per_step'], {k:round(v['ms_avg'],3) for k,v in d['roofline']['kernel_ms'].items()})
P
timeout 900 python tests/parity_sweep.py 5e6 1e6 Al,CH2,H2O 2>/dev/null | cut -c1-250
