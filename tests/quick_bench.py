import json,sys,subprocess,os
env=dict(os.environ)
for kv in sys.argv[1:]:
    k,v=kv.split('='); env[k]=v
out=subprocess.run([sys.executable,'bench.py','--steps','5','--warmup','3','--no-cpu-baseline'],capture_output=True,text=True,env=env).stdout.strip().splitlines()[-1]
d=json.loads(out)
print(sys.argv[1:], 'value %.3e ms_step %.3f fused %.3f xs %.3f sample %.3f e2e %.3e'%(d['value'],d['ms_per_step'],d['config']['ms_fused_xs_sample'],d['config']['ms_xs'],d['config']['ms_sample'],d['e2e']['value']))
