cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in x4 x5 x6 x8; do
  NCB200_LIB=$PWD/ncrystal_b200/libv/$v.so timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_$v.json 2> gpurun_out/r2f_bench_$v.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'Al value %.3e'%d['value'], 'xs %.3e'%d['config']['xs_per_s'], 'smp %.3e'%d['config']['samples_per_s'], 'e2e %.3e pageable %.3e ceil %.3e'%(d['e2e']['value'],d['e2e']['pageable']['value'],d['e2e']['copy_ceiling']['value']), 'classify', d['roofline']['kernel_ms']['k_sample_classify']['ms_avg'])
        for k,v in d['config']['other_configs'].items():
            print('    ',k, 'xs %.3e'%v.get('xs_per_s',0), 'smp %.3e'%v.get('samples_per_s',0), 'classify', (v.get('kernel_ms') or {}).get('k_sample_classify'), v.get('error'))
    except Exception as e: print(f,'ERR',e)
P
tail -3 gpurun_out/r2f_bench_x4.err
