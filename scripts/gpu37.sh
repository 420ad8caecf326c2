cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity_iso.py tests/test_gpu_parity_extra.py tests/test_gpu_api_semantics.py tests/test_gpu_fullsize_properties.py -x -q 2>&1 | tail -4
(time python bench.py --no-cpu-baseline) > gpurun_out/r2D_bench.json 2> gpurun_out/r2D_bench.err; tail -c 300 gpurun_out/r2D_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2D_bench.json').read().strip().splitlines()[-1])
print('value %.3e'%d['value'],'ms/step %.4f'%d['ms_per_step'],'smp %.3e'%(d['config']['samples_per_s']),'e2e %.3e'%d['e2e']['value'], {k:round(v['ms_avg'],3) for k,v in d['roofline']['kernel_ms'].items()})
for k,x in (d['config'].get('other_configs') or {}).items():
    print('    ',k, 'xs %.3e'%x.get('xs_per_s',0), 'smp %.3e'%x.get('samples_per_s',0), x.get('kernel_ms'))
P
