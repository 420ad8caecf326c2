cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2A_pytest.log; cat gpurun_out/r2A_pytest.log
(time python bench.py) > gpurun_out/r2A_bench.json 2> gpurun_out/r2A_bench.err; tail -c 300 gpurun_out/r2A_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2A_bench.json').read().strip().splitlines()[-1])
print('value %.3e'%d['value'],'xs %.3e smp %.3e'%(d['config']['xs_per_s'],d['config']['samples_per_s']),'e2e %.3e'%d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['kernel_ms'])
for k,x in (d['config'].get('other_configs') or {}).items():
    print('    ',k, 'xs %.3e'%x.get('xs_per_s',0), 'smp %.3e'%x.get('samples_per_s',0), x.get('neutrons_per_s'), x.get('error'))
print(d['config'].get('vdos_expansion')); print(d['config'].get('transport_step'))
P
timeout 600 python tests/mmc_bench.py 1e7 1e6 > gpurun_out/r2A_mmc_bench.jsonl 2> gpurun_out/r2A_mmc_bench.err; cut -c1-300 gpurun_out/r2A_mmc_bench.jsonl
