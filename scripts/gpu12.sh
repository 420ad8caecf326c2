cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 2400 python -m pytest tests -m gpu -x -q) > gpurun_out/r2l_pytest.log 2>&1; tail -8 gpurun_out/r2l_pytest.log
(time timeout 1200 python bench.py --steps 20 --warmup 5) > gpurun_out/r2l_bench_full.json 2> gpurun_out/r2l_bench_full.err
tail -5 gpurun_out/r2l_bench_full.err
NCB200_LIB=$PWD/ncrystal_b200/libv/nostream.so timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs > gpurun_out/r2l_bench_nostream.json 2> gpurun_out/r2l_bench_nostream.err
python - <<'P'
import json
for f in ('gpurun_out/r2l_bench_full.json','gpurun_out/r2l_bench_nostream.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f,'Al value %.3e'%d['value'], 'xs %.3e'%d['config']['xs_per_s'], 'smp %.3e'%d['config']['samples_per_s'], 'e2e %.3e pageable %.3e ceil %.3e'%(d['e2e']['value'],d['e2e']['pageable']['value'],d['e2e']['copy_ceiling']['value']), {k:round(v['ms_avg'],3) for k,v in d['roofline']['kernel_ms'].items()})
        for k,v in (d['config'].get('other_configs') or {}).items():
            print('    ',k, 'xs %.3e'%v.get('xs_per_s',0), 'smp %.3e'%v.get('samples_per_s',0), v.get('neutrons_per_s'), v.get('error'))
    except Exception as e: print(f,'ERR',e)
P
