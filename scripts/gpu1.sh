set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r2a_pytest.log 2>&1; tail -15 gpurun_out/r2a_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_cls.json 2> gpurun_out/r2a_bench_cls.err; tail -c 3000 gpurun_out/r2a_bench_cls.json
NCB200_CLS_MIN=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_old.json 2> gpurun_out/r2a_bench_old.err
for t in 3 8; do NCB200_CLS_TICKETS=$t timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_t$t.json 2>/dev/null; done
NCB200_SUBLAUNCH=2500000 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_sub25.json 2>/dev/null
NCB200_SUBLAUNCH=5000000 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_sub50.json 2>/dev/null
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2a_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'xs %.3e'%d['config']['xs_per_s'], 'smp %.3e'%d['config']['samples_per_s'], {k:round(v['ms_avg'],3) for k,v in d['roofline']['kernel_ms'].items()})
    except Exception as e: print(f,'ERR',e)
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sab_classes -s 3 -c 1 -o gpurun_out/r2a_sab_classes python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_c.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sample_classify -s 3 -c 1 -o gpurun_out/r2a_classify python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_d.log 2>&1
ls -la gpurun_out | tail
