cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tests/mmc_ktime.py Ge 1e6 2>&1 | tail -3 | tee gpurun_out/r2q_mmc_ktime.jsonl
timeout 600 python tests/mmc_ktime.py Al 1e7 2>&1 | tail -1 | tee -a gpurun_out/r2q_mmc_ktime.jsonl
(time python bench.py) > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -c 600 gpurun_out/r2q_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2q_bench.json').read().strip().splitlines()[-1])
print('value %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], d['config'].get('vdos_expansion'))
P
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2q_pytest.log; cat gpurun_out/r2q_pytest.log
