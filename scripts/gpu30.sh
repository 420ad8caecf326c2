cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
N=${NG:-8}
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5) > gpurun_out/r2F_bench_n$N.json 2> gpurun_out/r2F_bench_n$N.err; tail -c 400 gpurun_out/r2F_bench_n$N.err
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus $N --steps 3 --warmup 1) > gpurun_out/r2F_bench_ref_n$N.json 2> gpurun_out/r2F_bench_ref_n$N.err; tail -c 200 gpurun_out/r2F_bench_ref_n$N.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2F_bench_n$N.json').read().strip().splitlines()[-1])
print('value %.3e'%d['value'],'xs %.3e smp %.3e'%(d['config']['xs_per_s'],d['config']['samples_per_s']),'e2e %.3e'%d['e2e']['value'], d['e2e'].get('copy_ceiling'), d['roofline']['kernel'], d['roofline']['frac'])
for k,x in (d['config'].get('other_configs') or {}).items():
    print('    ',k, 'xs %.3e'%x.get('xs_per_s',0), 'smp %.3e'%x.get('samples_per_s',0), x.get('neutrons_per_s'), x.get('workload','')[:90], x.get('error'))
r=json.loads(open('gpurun_out/r2F_bench_ref_n$N.json').read().strip().splitlines()[-1])
print('reference arm value %.3e'%r['value'], r.get('cpu_baseline'))
P
