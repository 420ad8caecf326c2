cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NCB200_LIB=$GRAFT_REPO_ROOT/ncrystal_b200/lib/libncrystal_b200_tuning.so
for ov in 0 8 7 6 5 4 3; do
  NCB200_FG_OVERLAP=$ov timeout 300 python bench.py --no-cpu-baseline --no-other-configs --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=r['roofline']['kernel_ms']
print('overlap', $ov, 'value %.4g ms/step %.4f' % (r['value'], r['ms_per_step']), {a:round(b['ms_avg'],3) for a,b in k.items()})
" | tee -a gpurun_out/r2P_overlap.txt
done
