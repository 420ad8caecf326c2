cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in v0 v1 v2 v3 v4; do
  NCB200_LIB=$PWD/ncrystal_b200/libv/$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs > gpurun_out/r2d_bench_$v.json 2> gpurun_out/r2d_bench_$v.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2d_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'smp %.3e'%d['config']['samples_per_s'], {k:round(v['ms_avg'],3) for k,v in d['roofline']['kernel_ms'].items() if 'refill' in k or 'classify' in k})
    except Exception as e: print(f,'ERR',e)
P
