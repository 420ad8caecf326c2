set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r2b_pytest.log 2>&1; tail -15 gpurun_out/r2b_pytest.log
for mb in 8 7 6; do NCB200_SAB_MINB=$mb timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs > gpurun_out/r2b_bench_mb$mb.json 2> gpurun_out/r2b_bench_mb$mb.err; done
(time timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/r2b_bench_full.json 2> gpurun_out/r2b_bench_full.err
tail -5 gpurun_out/r2b_bench_full.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'xs %.3e'%d['config']['xs_per_s'], 'smp %.3e'%d['config']['samples_per_s'], 'e2e %.3e'%d['e2e']['value'], {k:round(v['ms_avg'],3) for k,v in d['roofline']['kernel_ms'].items()})
    except Exception as e: print(f,'ERR',e)
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sample_sab_refill -s 6 -c 1 -o gpurun_out/r2b_refill python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2b_ncu_c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_xs_iso -s 2 -c 1 -o gpurun_out/r2b_xs python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2b_ncu_x.log 2>&1
ls -la gpurun_out | tail -12
