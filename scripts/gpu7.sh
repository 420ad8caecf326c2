cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 2400 python -m pytest tests -m gpu -x -q) > gpurun_out/r2g_pytest.log 2>&1; tail -12 gpurun_out/r2g_pytest.log
(time timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/r2g_bench_full.json 2> gpurun_out/r2g_bench_full.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2g_bench_full.json').read().strip().splitlines()[-1])
print('Al value %.3e'%d['value'], 'xs %.3e'%d['config']['xs_per_s'], 'smp %.3e'%d['config']['samples_per_s'], 'e2e %.3e pageable %.3e ceil %.3e'%(d['e2e']['value'],d['e2e']['pageable']['value'],d['e2e']['copy_ceiling']['value']), {k:round(v['ms_avg'],3) for k,v in d['roofline']['kernel_ms'].items()})
for k,v in d['config']['other_configs'].items():
    print('    ',k, 'xs %.3e'%v.get('xs_per_s',0), 'smp %.3e'%v.get('samples_per_s',0), 'classify', (v.get('kernel_ms') or {}).get('k_sample_classify'), v.get('error'))
P
for t in 3 7 15; do NCB200_COPY_THREADS=$t python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('copy threads $t: pageable %.3e pinned %.3e'%(d['e2e']['pageable']['value'], d['e2e']['value']))"; done
