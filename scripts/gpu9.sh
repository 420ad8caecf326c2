cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(time timeout 2400 python -m pytest tests -m gpu -x -q) > gpurun_out/r2i_pytest.log 2>&1; tail -6 gpurun_out/r2i_pytest.log
(time timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/r2i_bench_full.json 2> gpurun_out/r2i_bench_full.err
tail -3 gpurun_out/r2i_bench_full.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2i_bench_full.json').read().strip().splitlines()[-1])
print('Al value %.3e'%d['value'], 'xs %.3e'%d['config']['xs_per_s'], 'smp %.3e'%d['config']['samples_per_s'], 'e2e %.3e pageable %.3e ceil %.3e'%(d['e2e']['value'],d['e2e']['pageable']['value'],d['e2e']['copy_ceiling']['value']), {k:round(v['ms_avg'],3) for k,v in d['roofline']['kernel_ms'].items()})
print('roofline', d['roofline']['kernel'], d['roofline']['frac'], 'cpu', d.get('cpu_baseline'))
for k,v in d['config']['other_configs'].items():
    print('    ',k, 'xs %.3e'%v.get('xs_per_s',0), 'smp %.3e'%v.get('samples_per_s',0), v.get('kernel_ms'), v.get('error'))
print('transport', d['config']['transport_step'])
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2i_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2i_ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sample_|k_fg_|k_tally" -s 18 -c 9 -o gpurun_out/r2i_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2i_ncu_s.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_xs_iso -s 1 -c 1 -o gpurun_out/r2i_xs python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2i_ncu_x.log 2>&1
for r in r2i_step r2i_xs; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null; done
ls -la gpurun_out | tail -12
