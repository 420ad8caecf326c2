cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
gcc -std=c99 -D_POSIX_C_SOURCE=200809L -pedantic -Wall -Werror -Iinclude tests/capi_multigpu.c -o /tmp/capi_multigpu -Lncrystal_b200/lib -lncrystal_b200 -lm -Wl,-rpath,$PWD/ncrystal_b200/lib
for nd in 1 2; do /tmp/capi_multigpu "Al_sg225.ncmat;temp=293.15K" $nd 10000000 | tee -a gpurun_out/r2n_capi_multigpu.jsonl; done
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5) > gpurun_out/r2n_bench_n2.json 2> gpurun_out/r2n_bench_n2.err; tail -c 800 gpurun_out/r2n_bench_n2.err
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1) > gpurun_out/r2n_bench_ref_n2.json 2> gpurun_out/r2n_bench_ref_n2.err; tail -c 300 gpurun_out/r2n_bench_ref_n2.err; cat gpurun_out/r2n_bench_ref_n2.json | cut -c1-400
python - <<'P'
import json
for f in ('gpurun_out/r2n_bench_n2.json',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f,'value %.3e'%d['value'],'xs %.3e smp %.3e'%(d['config']['xs_per_s'],d['config']['samples_per_s']),'e2e',d['e2e']['value'], d['e2e'].get('copy_ceiling'), d['roofline']['kernel'], d['roofline']['frac'], d.get('cpu_baseline',{}).get('value'))
        for k,x in (d['config'].get('other_configs') or {}).items():
            print('    ',k, 'xs %.3e'%x.get('xs_per_s',0), 'smp %.3e'%x.get('samples_per_s',0), x.get('neutrons_per_s'), x.get('error'))
    except Exception as e: print(f,'ERR',e)
P
