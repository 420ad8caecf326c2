cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r2e_pytest.log 2>&1; tail -4 gpurun_out/r2e_pytest.log
(time timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/r2e_bench_full.json 2> gpurun_out/r2e_bench_full.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2e_bench_full.json').read().strip().splitlines()[-1])
print('Al value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'xs %.3e'%d['config']['xs_per_s'], 'smp %.3e'%d['config']['samples_per_s'], 'e2e %.3e'%d['e2e']['value'], {k:round(v['ms_avg'],3) for k,v in d['roofline']['kernel_ms'].items()})
for k,v in d['config']['other_configs'].items():
    print(k, 'xs %.3e'%v.get('xs_per_s',0), 'smp %.3e'%v.get('samples_per_s',0), v.get('kernel_ms'), v.get('error'))
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_xs_iso -s 2 -c 1 -o gpurun_out/r2e_xs python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2e_ncu_x.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sample_classify -s 3 -c 1 -o gpurun_out/r2e_classify python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2e_ncu_d.log 2>&1
